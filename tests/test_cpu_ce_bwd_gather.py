"""CPU replay of ce_bwd_gather_kernel's index arithmetic (csrc/loss.cu): the separable, atomic-free pull-back of the
cross-entropy gradient through the bilinear up-sampling — tap tables per tile, the per-column label ranges, the
8-row chunks and the per-chunk head-row window — restated in numpy and checked against torch autograd of
F.interpolate + F.cross_entropy.  The device kernel itself is checked on the GPU (tests/test_gpu_golden.py::
test_loss_golden, tests/test_gpu_baseline_config.py::test_baseline_config_criteria)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

TILE, ROWS = 32, 8


def taps(n_in, n_out, align, o):
    """resample.cuh bilinear_taps in fp32."""
    f = np.float32
    if align:
        scale = f(n_in - 1) / f(n_out - 1) if n_out > 1 else f(0)
        src = scale * f(o)
    else:
        scale = f(n_in) / f(n_out)
        src = max(f(0), scale * (f(o) + f(0.5)) - f(0.5))
    i0 = min(int(src), n_in - 1)
    i1 = i0 + (1 if i0 < n_in - 1 else 0)
    l1 = f(src) - f(i0)
    return i0, i1, f(1) - l1, l1, scale


def replay(G, h, w, align):
    """G [n, lh, lw, c] = d loss / d up-sampled logits -> d logits [n, c, h, w], the kernel's way."""
    n, lh, lw, c = G.shape
    out = np.zeros((n, c, h, w), dtype=np.float64)
    scale_h, scale_w = taps(h, lh, align, 0)[4], taps(w, lw, align, 0)[4]
    HR, WR = int((TILE - 1) * scale_h) + 3, int((TILE - 1) * scale_w) + 3
    for b in range(n):
        for ty0 in range(0, lh, TILE):
            for tx0 in range(0, lw, TILE):
                hy0, hx0 = taps(h, lh, align, ty0)[0], taps(w, lw, align, tx0)[0]
                xt = [taps(w, lw, align, min(tx0 + i, lw - 1)) for i in range(TILE)]
                yt = [taps(h, lh, align, min(ty0 + i, lh - 1)) for i in range(TILE)]
                xw0, xw1 = [t[0] - hx0 for t in xt], [t[1] - hx0 for t in xt]
                yh0, yh1 = [t[0] - hy0 for t in yt], [t[1] - hy0 for t in yt]
                assert max(xw1) < WR and max(yh1) < HR
                xs, xe = [TILE] * WR, [0] * WR
                for wq in range(WR):
                    for x in range(TILE):
                        if xw0[x] == wq or xw1[x] == wq:
                            xs[wq], xe[wq] = min(xs[wq], x), max(xe[wq], x + 1)
                acc = np.zeros((c, HR, WR))
                for chunk in range(TILE // ROWS):
                    Gt = np.zeros((ROWS, c, TILE))
                    for r in range(ROWS):
                        y = ty0 + chunk * ROWS + r
                        for x in range(TILE):
                            if y < lh and tx0 + x < lw:
                                Gt[r, :, x] = G[b, y, tx0 + x]
                    R = np.zeros((ROWS, c, WR))
                    for wq in range(WR):
                        for x in range(xs[wq], xe[wq]):
                            wt = (xt[x][2] if xw0[x] == wq else 0) + (xt[x][3] if xw1[x] == wq else 0)
                            R[:, :, wq] += wt * Gt[:, :, x]
                    hlo, hhi = yh0[chunk * ROWS], yh1[chunk * ROWS + ROWS - 1]
                    for hh in range(hlo, hhi + 1):
                        for r in range(ROWS):
                            ry = chunk * ROWS + r
                            wt = (yt[ry][2] if yh0[ry] == hh else 0) + (yt[ry][3] if yh1[ry] == hh else 0)
                            acc[:, hh, :] += wt * R[r]
                for r in range(HR):
                    for q in range(WR):
                        if hy0 + r < h and hx0 + q < w:
                            out[b, :, hy0 + r, hx0 + q] += acc[:, r, q]
    return out


@pytest.mark.parametrize("h,w,lh,lw,align", [(12, 12, 48, 48, True), (10, 7, 40, 28, True), (9, 11, 36, 44, False),
                                              (16, 16, 16, 16, True), (5, 6, 37, 50, False)])
def test_gather_form_equals_autograd(h, w, lh, lw, align):
    gen = torch.Generator().manual_seed(h * 100 + lw)
    logits = torch.randn(2, 3, h, w, generator=gen, dtype=torch.float64).requires_grad_(True)
    target = torch.randint(0, 3, (2, lh, lw), generator=gen)
    target[:, :2] = 255
    up = F.interpolate(logits, size=(lh, lw), mode="bilinear", align_corners=align)
    up.retain_grad()
    F.cross_entropy(up, target, ignore_index=255, reduction="sum").backward()
    G = up.grad.permute(0, 2, 3, 1).numpy()
    got = replay(G, h, w, align)
    want = logits.grad.numpy()
    assert np.abs(got - want).max() < 1e-5 * max(1.0, np.abs(want).max())
