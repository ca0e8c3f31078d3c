"""ORACLE — test infrastructure only.  numpy restatement of the reference's label synthesis
(/root/reference/dataset/target_generation.py), per image like the reference:

  gen_single_gaussian_map / gen_pose_target   :146-168 / :94-121
  generate_edge                               :210-239 (cv2.dilate with a 3x3 rectangle restated as a max filter with a
                                              constant border) + dataset/data_loader.py:284 (255 where the label is 255)
  flip_parsing                                :44-56 (cv2.flip + left/right relabel)
  flip_joints                                 :8-24

Pinned against the reference's own functions imported from /root/reference (tests/test_oracle_labels.py, which also
writes / checks tests/golden/labels_golden.npz)."""
import numpy as np


def gen_single_gaussian_map(center, stride, grid_x, grid_y, sigma):
    g = np.zeros((grid_y, grid_x))
    start = stride / 2.0 - 0.5
    max_dist = np.ceil(np.sqrt(4.6052 * sigma * sigma * 2.0))
    sx = int(max(0, np.floor((center[0] - max_dist - start) / stride)))
    ex = int(min(grid_x, np.ceil((center[0] + max_dist - start) / stride)))
    sy = int(max(0, np.floor((center[1] - max_dist - start) / stride)))
    ey = int(min(grid_y, np.ceil((center[1] + max_dist - start) / stride)))
    if ex <= sx or ey <= sy:
        return g
    xs = start + np.arange(sx, ex) * stride
    ys = start + np.arange(sy, ey) * stride
    d2 = ((xs - center[0]) * (xs - center[0]))[None, :] + ((ys - center[1]) * (ys - center[1]))[:, None]
    e = d2 / 2.0 / sigma / sigma
    v = np.where(e > 4.6052, 0.0, np.exp(-e))
    g[sy:ey, sx:ex] = np.minimum(v, 1.0)
    return g


def gen_pose_target(joints, visibility, stride=8, grid_x=46, grid_y=46, sigma=7, aux=False):
    def maps(sig):
        n = joints.shape[0]
        m = np.zeros((n + 1, grid_y, grid_x))
        for ji in range(n):
            if visibility[ji]:
                m[ji] = gen_single_gaussian_map(joints[ji, :], stride, grid_x, grid_y, sig)
        m[n] = 1 - m.max(0)
        return m
    return maps(sigma), (maps(2 * sigma) if aux else None)


def generate_edge(label, edge_width=3, mark_ignore=True):
    h, w = label.shape
    edge = np.zeros(label.shape)
    e = edge[1:h, :]
    e[(label[1:h, :] != label[:h - 1, :]) & (label[1:h, :] != 255) & (label[:h - 1, :] != 255)] = 1
    e = edge[:, :w - 1]
    e[(label[:, :w - 1] != label[:, 1:w]) & (label[:, :w - 1] != 255) & (label[:, 1:w] != 255)] = 1
    e = edge[:h - 1, :w - 1]
    e[(label[:h - 1, :w - 1] != label[1:h, 1:w]) & (label[:h - 1, :w - 1] != 255) & (label[1:h, 1:w] != 255)] = 1
    e = edge[:h - 1, 1:w]
    e[(label[:h - 1, 1:w] != label[1:h, :w - 1]) & (label[:h - 1, 1:w] != 255) & (label[1:h, :w - 1] != 255)] = 1
    r = edge_width // 2
    pad = np.zeros((h + 2 * r, w + 2 * r))
    pad[r:r + h, r:r + w] = edge
    out = np.zeros_like(edge)
    for dy in range(edge_width):
        for dx in range(edge_width):
            out = np.maximum(out, pad[dy:dy + h, dx:dx + w])
    if mark_ignore:
        out[label == 255] = 255
    return out


def flip_parsing(label):
    out = label[:, ::-1].copy()
    for r, l in ((15, 14), (17, 16), (19, 18)):
        rp, lp = out == r, out == l
        out[rp], out[lp] = l, r
    return out


def flip_joints(joints, im_w, r_joint=(0, 1, 2, 10, 11, 12), l_joint=(3, 4, 5, 13, 14, 15)):
    f = joints.copy()
    f[:, 0] = im_w - 1 - f[:, 0]
    for r, l in zip(r_joint, l_joint):
        f[[r, l]] = f[[l, r]]
    return f
