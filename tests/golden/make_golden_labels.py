"""Generates tests/golden/labels_golden.npz FROM THE REFERENCE ITSELF (/root/reference/dataset/target_generation.py):
Gaussian pose targets, edge maps, flipped parsing labels and flipped joints for seeded synthetic inputs.

  python tests/golden/make_golden_labels.py        (build container only; needs cv2)
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_module():
    spec = importlib.util.spec_from_file_location("ref_target_generation", "/root/reference/dataset/target_generation.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def inputs(seed=0, b=3, nj=16, size=96, stride=4):
    rng = np.random.RandomState(seed)
    joints = rng.uniform(-20, size * stride + 20, size=(b, nj, 2))       # some joints outside the crop
    joints[0, 0] = [0.0, 0.0]
    joints[0, 1] = [size * stride - 1.0, size * stride - 1.0]
    vis = (rng.uniform(size=(b, nj)) > 0.2).astype(np.int32)
    h, w = 64, 80
    label = rng.randint(0, 20, size=(b, h // 8, w // 8)).repeat(8, axis=1).repeat(8, axis=2).astype(np.uint8)
    noise = rng.uniform(size=label.shape) < 0.02
    label[noise] = rng.randint(0, 20, size=int(noise.sum()))
    label[:, :3, :] = 255
    label[:, 20:24, 30:50] = 255
    return joints, vis, label


def main():
    T = reference_module()
    joints, vis, label = inputs()
    out = {"joints": joints, "vis": vis, "label": label, "stride": np.array(4), "grid": np.array(96), "sigma": np.array(7)}
    maps, aux = [], []
    for b in range(joints.shape[0]):
        m, a = T.gen_pose_target(joints[b], vis[b], 4, 96, 96, 7, aux=True)
        maps.append(m)
        aux.append(a)
    out["pose"], out["pose_aux"] = np.stack(maps), np.stack(aux)                     # float64, as the reference returns
    edges, flips = [], []
    for b in range(label.shape[0]):
        e = T.generate_edge(label[b])
        e[label[b] == 255] = 255                                                   # dataset/data_loader.py:284
        edges.append(e)
        flips.append(T.gen_parsing_target(label[b], flip_param=True, stride=1))
    out["edge"], out["flip"] = np.stack(edges), np.stack(flips)
    out["flip_joints"] = np.stack([T.flip_joints(joints[b], 384) for b in range(joints.shape[0])])
    np.savez_compressed(os.path.join(HERE, "labels_golden.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
