"""Label synthesis on the GPU — the batched device side of the reference's dataset/target_generation.py (SURVEY.md §8f
N3).  The reference builds these targets per image inside DataLoader workers (dataset/data_loader.py:239-285: Gaussian
heat maps, edge map, flip relabel) with numpy / cv2 loops; at > 1000 img/s per GPU that host path cannot feed the
training step.  Function names and argument meaning follow the reference; inputs and outputs carry a leading batch axis
and live on the device."""
import numpy as np
import torch

from .._lib import call, fptr, i32, f64, stream


def _dev(t, dtype):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    return t.to(device="cuda", dtype=dtype).contiguous()


def gen_pose_target(joints, visibility, stride=8, grid_x=46, grid_y=46, sigma=7, aux=False):
    """target_generation.py:94-121.  joints [B, J, 2] (x, y), visibility [B, J] -> (maps, maps_aux | None), each fp32
    [B, J + 1, grid_y, grid_x] with the background map last (the training loop drops it, core/function.py:78-79)."""
    j = _dev(joints, torch.float64)
    v = _dev(visibility, torch.int32)
    if j.dim() != 3 or j.shape[2] != 2 or v.shape != j.shape[:2]:
        raise AssertionError("joints must be [B, J, 2] and visibility [B, J]")
    b, nj = v.shape

    def maps(sig):
        out = torch.empty((b, nj + 1, int(grid_y), int(grid_x)), dtype=torch.float32, device=j.device)
        call("npp_pose_target", fptr(j), fptr(v), i32(b), i32(nj), f64(stride), i32(grid_x), i32(grid_y), f64(sig), fptr(out),
             stream())
        return out

    return maps(sigma), (maps(2 * sigma) if aux else None)


def generate_edge(label, edge_width=3):
    """target_generation.py:210-239 followed by data_loader.py:284 (`parsing_edge[parsing_target == 255] = 255`):
    label int [B, H, W] -> int64 [B, H, W] in {0, 1, 255}."""
    lab = _dev(label, torch.int64)
    if lab.dim() != 3:
        raise AssertionError("label must be [B, H, W]")
    b, h, w = lab.shape
    out = torch.empty_like(lab)
    call("npp_edge_label", fptr(lab), i32(b), i32(h), i32(w), i32(edge_width), fptr(out), stream())
    return out


def flip_parsing(label):
    """gen_parsing_target's flip branch (target_generation.py:44-56): mirror + swap left/right part labels."""
    lab = _dev(label, torch.int64)
    b, h, w = lab.shape
    out = torch.empty_like(lab)
    call("npp_flip_parsing", fptr(lab), i32(b), i32(h), i32(w), fptr(out), stream())
    return out


def flip_joints(joints, im_w, r_joint=(0, 1, 2, 10, 11, 12), l_joint=(3, 4, 5, 13, 14, 15)):
    """target_generation.py:8-24 on a [B, J, 2] (or [J, 2]) array: x -> im_w - 1 - x, then swap left / right joints.
    Host arithmetic (a few hundred numbers per batch)."""
    j = np.array(joints, dtype=np.float64, copy=True)
    j[..., 0] = im_w - 1 - j[..., 0]
    r, l = list(r_joint), list(l_joint)
    tmp = j[..., r, :].copy()
    j[..., r, :] = j[..., l, :]
    j[..., l, :] = tmp
    return j
