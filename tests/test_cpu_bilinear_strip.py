"""CPU replay of the index arithmetic of the strip-form bilinear up-sampling forward (csrc/resample.cu
`bilinear_fwd_strip_t`) and of the batched first pass of the separable backward (`bilinear_bwd_sep_w_batched`): the
kernels fetch a FIXED window of input / output columns per thread, so every tap the reference arithmetic (ATen
area_pixel_compute_source_index in fp32, resample.cuh `bilinear_taps`) can produce must fall inside that window."""
import numpy as np
import pytest

F = np.float32


def axis_scale(n_in, n_out, align, scale_factor=None):
    if align:
        return F(n_in - 1) / F(n_out - 1) if n_out > 1 else F(0)
    return F(1.0 / scale_factor) if scale_factor else F(n_in) / F(n_out)


def taps(n_in, scale, align, o):
    """bilinear_taps: (i0, i1, l0, l1) in fp32 arithmetic."""
    if align:
        src = F(scale * F(o))
    else:
        src = F(F(scale * F(F(o) + F(0.5))) - F(0.5))
        if src < 0:
            src = F(0)
    i0 = min(int(src), n_in - 1)
    i1 = i0 + (1 if i0 < n_in - 1 else 0)
    l1 = F(src - F(i0))
    return i0, i1, F(F(1) - l1), l1


CASES = [(n_in, n_in * k, align, float(k)) for n_in in (6, 12, 24, 48, 96, 128, 17, 23) for k in (2, 4, 8)
         for align in (True, False)]
CASES += [(24, 96, True, 4.0), (48, 96, True, 2.0), (12, 96, True, 8.0), (96, 384, True, 4.0), (128, 512, True, 4.0)]


@pytest.mark.parametrize("n_in,n_out,align,sf", CASES)
def test_strip_forward_window_covers_every_tap(n_in, n_out, align, sf):
    scale = axis_scale(n_in, n_out, align, sf)
    assert 0 < scale <= 0.5            # the dispatch condition of the strip kernel
    KW = 4
    for wo0 in range(0, n_out, KW):
        cols = [taps(n_in, scale, align, min(wo0 + k, n_out - 1)) for k in range(KW)]
        base = cols[0][0]
        for k, (w0, w1, l0, l1) in enumerate(cols):
            if wo0 + k >= n_out:
                break
            assert 0 <= w0 - base <= 3 and 0 <= w1 - base <= 3, (wo0, k, base, w0, w1)
            # the clamped fetch min(base + j, n_in - 1) returns the right column for both taps
            assert min(base + (w0 - base), n_in - 1) == w0 and min(base + (w1 - base), n_in - 1) == w1


@pytest.mark.parametrize("n_in,n_out,align,sf", [c for c in CASES if c[3] in (2.0, 4.0)])
def test_batched_backward_window_covers_every_contribution(n_in, n_out, align, sf):
    scale = axis_scale(n_in, n_out, align, sf)
    maxc = 8 if scale >= 0.45 else 13
    assert (0.45 <= scale <= 1.0) or (0.2 <= scale < 0.45)
    inv, off = F(1) / scale, F(0.0 if align else 0.5)
    contrib = {w: [] for w in range(n_in)}
    for wo in range(n_out):
        w0, w1, l0, l1 = taps(n_in, scale, align, wo)
        if l0 != 0:
            contrib[w0].append(wo)
        if l1 != 0 and w1 != w0:
            contrib[w1].append(wo)
    for w in range(n_in):
        lo = int(np.floor(F(F(F(w) - F(1) + off) * inv) - off)) - 1
        lo = max(lo, 0)
        if lo > n_out - maxc:
            lo = max(n_out - maxc, 0)
        for wo in contrib[w]:
            assert lo <= wo < lo + maxc, (w, wo, lo, maxc)
