"""CPU restatement (numpy) of the LIP pose post-process of `validate_sync` — SURVEY.md §8(f) row N1.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and later by bench.py's CPU leg); the product path never calls it.

What the reference does per image and joint (core/function.py:962-986, same code at :806-830 and :1140-1170):

    heatmap  = cv2.resize(pred[num, ji], (W, H), INTER_LINEAR)                      # 96x96 -> 384x384
    flipped  = cv2.flip(cv2.resize(flip_pred[num, flipped_poseidx[ji]], (W, H), INTER_LINEAR), 1)
    heatmap  = scipy.ndimage.gaussian_filter((heatmap + flipped) * 0.5, sigma=3)
    (py, px) = np.unravel_index(heatmap.argmax(), heatmap.shape)
    x = (px - crop[0, 2] + crop[0, 0]) / scale ;  y = (py - crop[0, 3] + crop[0, 1]) / scale
    pose[num, ji] = (x, y, heatmap[py, px])

and `save_hpe_results_to_lip_format` (utils/utils.py:270-289) writes int(x), int(y) in LIP joint order.

The arithmetic lives in two third-party libraries that are not vendored under /root/reference
(requirements.txt: opencv-python, scipy); their published algorithms are restated here and pinned against the
libraries themselves (tests/golden/make_golden_pose.py -> tests/golden/pose_post_golden.npz, and live in
tests/test_oracle_pose_post.py when cv2 / scipy import):

  * cv2.resize(INTER_LINEAR) on float32: half-pixel centres, source index clamped to the image
    (imgproc/resize.cpp: fx = (dx + 0.5) * scale - 0.5; sx = floor(fx); sx < 0 -> (0, weight 0);
    sx >= w - 1 -> (w - 1, weight 0)), horizontal pass then vertical pass in float32.
  * scipy.ndimage.gaussian_filter(sigma=3): separable, axis 0 then axis 1, radius int(4 * sigma + 0.5) = 12,
    weights exp(-x^2 / (2 sigma^2)) normalised in float64, mode 'reflect' (d c b a | a b c d | d c b a),
    accumulation in float64, each pass rounded to the input dtype (float32).
"""
import numpy as np

FLIPPED_POSEIDX = [0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 14, 15]   # core/function.py:908
IDX_MAP_TO_LIP = [10, 9, 8, 11, 12, 13, 15, 14, 1, 0, 4, 3, 2, 5, 6, 7]    # utils/utils.py:279


def _linear_taps(dst, src):
    """cv2 INTER_LINEAR source taps of one axis: (index0, index1, weight1 as float32)."""
    scale = float(src) / float(dst)
    i0 = np.empty(dst, dtype=np.int64)
    w1 = np.empty(dst, dtype=np.float32)
    for d in range(dst):
        f = (d + 0.5) * scale - 0.5
        s = int(np.floor(f))
        f -= s
        if s < 0:
            s, f = 0, 0.0
        if s >= src - 1:
            s, f = src - 1, 0.0
        i0[d] = s
        w1[d] = np.float32(f)
    i1 = np.minimum(i0 + 1, src - 1)
    return i0, i1, w1


def resize_bilinear_cv2(img, out_h, out_w):
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR) for a 2-D float32 array."""
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape
    x0, x1, ax = _linear_taps(out_w, w)
    y0, y1, ay = _linear_taps(out_h, h)
    one = np.float32(1.0)
    rows = img[:, x0] * (one - ax)[None, :] + img[:, x1] * ax[None, :]          # horizontal pass, float32
    rows = rows.astype(np.float32)
    out = rows[y0, :] * (one - ay)[:, None] + rows[y1, :] * ay[:, None]         # vertical pass, float32
    return out.astype(np.float32)


def _gaussian_kernel1d(sigma, truncate=4.0):
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return k / k.sum(), radius


def _correlate1d_reflect(a, k, radius, axis):
    """scipy.ndimage.correlate1d(mode='reflect') along `axis`, float64 accumulation, result in a's dtype."""
    a = np.asarray(a)
    pad = [(0, 0)] * a.ndim
    pad[axis] = (radius, radius)
    ap = np.pad(a.astype(np.float64), pad, mode="symmetric")   # numpy 'symmetric' == scipy 'reflect'
    out = np.zeros(a.shape, dtype=np.float64)
    n = a.shape[axis]
    for j in range(2 * radius + 1):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(j, j + n)
        out += k[j] * ap[tuple(sl)]
    return out.astype(a.dtype)


def gaussian_filter_reflect(img, sigma=3.0, truncate=4.0):
    """scipy.ndimage.gaussian_filter(img, sigma) for a 2-D float32 array (default mode / truncate)."""
    k, radius = _gaussian_kernel1d(sigma, truncate)
    out = _correlate1d_reflect(np.asarray(img, dtype=np.float32), k, radius, 0)
    return _correlate1d_reflect(out, k, radius, 1)


def merged_heatmap(pred, flip_pred, num, ji, out_h, out_w):
    """The filtered, flip-averaged, up-sampled heat map of (image num, joint ji) — function.py:973-981."""
    hm = resize_bilinear_cv2(pred[num, ji], out_h, out_w)
    fl = resize_bilinear_cv2(flip_pred[num, FLIPPED_POSEIDX[ji]], out_h, out_w)[:, ::-1]
    hm = (hm + fl).astype(np.float32) * np.float32(0.5)
    return gaussian_filter_reflect(hm, sigma=3.0)


def pose_postprocess(pred, flip_pred, size, crop_param, scale):
    """pred, flip_pred: [N, 16, h, w] float32 heat maps (the network output and the output for the mirrored image);
    size = (H, W) of the network input; crop_param: [N, >=1, 4]; scale: [N].  Returns pose [N, 16, 3] float64 =
    (x, y, peak value) in original-image coordinates (function.py:969-986)."""
    pred = np.asarray(pred, dtype=np.float32)
    flip_pred = np.asarray(flip_pred, dtype=np.float32)
    n, nj = pred.shape[:2]
    out_h, out_w = int(size[0]), int(size[1])
    pose = np.zeros((n, nj, 3))
    for num in range(n):
        base_scale = float(scale[num])
        cp = np.asarray(crop_param[num], dtype=np.float64)
        for ji in range(nj):
            hm = merged_heatmap(pred, flip_pred, num, ji, out_h, out_w)
            py, px = np.unravel_index(hm.argmax(), hm.shape)
            pose[num, ji, 0] = (px - cp[0, 2] + cp[0, 0]) / base_scale
            pose[num, ji, 1] = (py - cp[0, 3] + cp[0, 1]) / base_scale
            pose[num, ji, 2] = hm[py, px]
    return pose


def lip_csv_rows(pose):
    """The integer coordinates save_hpe_results_to_lip_format writes (utils/utils.py:278-286): per image 32 ints,
    (x, y) of the joints in LIP order, truncated toward zero by int()."""
    pose = np.asarray(pose)
    rows = []
    for ii in range(pose.shape[0]):
        r = []
        for j in IDX_MAP_TO_LIP:
            r.append(int(pose[ii, j, 0]))
            r.append(int(pose[ii, j, 1]))
        rows.append(r)
    return np.asarray(rows, dtype=np.int64)
