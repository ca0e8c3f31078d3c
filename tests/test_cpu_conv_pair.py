"""CPU (float64): the index mapping behind the pixel-pair form of the 32 -> 32 channel 3x3 convolutions
(csrc/conv_tcgen05.cu pair_weight / fold_pair_wgrad_kernel, functional._ConvPairFn) — restated with torch ops and
checked against torch's own conv2d: forward, data gradient, weight-gradient fold and BatchNorm-sum fold."""
import torch


def pair_weight(w):
    """[32,32,3,3] -> [64,64,3,3]: rows po*32+co, cols pi*32+ci, horizontal tap ds+1 carries w[.., s = 2*ds+pi-po+1]."""
    wp = torch.zeros(64, 64, 3, 3, dtype=w.dtype)
    for po in range(2):
        for pi in range(2):
            for dsi in range(3):
                s = 2 * (dsi - 1) + pi - po + 1
                if 0 <= s <= 2:
                    wp[po * 32:(po + 1) * 32, pi * 32:(pi + 1) * 32, :, dsi] = w[:, :, :, s]
    return wp


def fold_pair_wgrad(dwp):
    dw = torch.zeros(32, 32, 3, 3, dtype=dwp.dtype)
    for s in range(3):
        for po in range(2):
            for pi in range(2):
                t2 = s - 1 - pi + po
                if t2 & 1 or not -1 <= t2 // 2 <= 1:
                    continue
                dw[:, :, :, s] += dwp[po * 32:(po + 1) * 32, pi * 32:(pi + 1) * 32, :, t2 // 2 + 1]
    return dw


def to_pair(t):   # NCHW view of "[N,H,W,32] read as [N,H,W/2,64]"
    n, c, h, w = t.shape
    return t.view(n, c, h, w // 2, 2).permute(0, 4, 1, 2, 3).reshape(n, 2 * c, h, w // 2)


def from_pair(t):
    n, c2, h, w2 = t.shape
    return t.view(n, 2, c2 // 2, h, w2).permute(0, 2, 3, 4, 1).reshape(n, c2 // 2, h, 2 * w2)


def test_pair_mapping_equals_conv2d():
    torch.manual_seed(0)
    w = torch.randn(32, 32, 3, 3, dtype=torch.float64)
    x = torch.randn(2, 32, 7, 10, dtype=torch.float64)
    gy = torch.randn(2, 32, 7, 10, dtype=torch.float64)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y = torch.nn.functional.conv2d(xr, wr, padding=1)
    y.backward(gy)
    xp, wp = to_pair(x).clone().requires_grad_(True), pair_weight(w).clone().requires_grad_(True)
    yp = torch.nn.functional.conv2d(xp, wp, padding=1)
    yp.backward(to_pair(gy))
    assert (from_pair(yp) - y).abs().max() < 1e-12
    assert (from_pair(xp.grad) - xr.grad).abs().max() < 1e-12
    assert (fold_pair_wgrad(wp.grad) - wr.grad).abs().max() < 1e-11
    sums = torch.cat([yp.sum((0, 2, 3)), (yp * yp).sum((0, 2, 3))]).detach()          # what the epilogue produces
    folded = sums.view(2, 2, 32).sum(1).reshape(64)
    want = torch.cat([y.sum((0, 2, 3)), (y * y).sum((0, 2, 3))]).detach()
    assert (folded - want).abs().max() < 1e-9
