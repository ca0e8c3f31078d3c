"""CPU: the index arithmetic of pack_weights_tiles_kernel (csrc/conv_tcgen05.cu, NPP_PACK_TILES=1 candidate) replayed
in numpy — tile decomposition, shared-memory row stride 32*taps+1, both output layouts, zero padding — against the
definition of the packed layouts (w[co][tap][ci], wt[ci][tap][co] of the OIHW master, channels padded to 8)."""
import numpy as np
import pytest


def _pad8(c):
    return (c + 7) // 8 * 8


def _replay(w32):
    cout, cin, taps = w32.shape
    cop, cip = _pad8(cout), _pad8(cin)
    flat = w32.reshape(-1)
    w = np.full(cop * taps * cip, 7.0, np.float32)
    wt = np.full(cop * taps * cip, 7.0, np.float32)
    ci_tiles = (cip + 31) // 32
    ld = 32 * taps + 1
    for tix in range(((cop + 31) // 32) * ci_tiles):
        co0, ci0 = (tix // ci_tiles) * 32, (tix % ci_tiles) * 32
        nci = min(cin - ci0, 32)
        tile = np.zeros(32 * ld, np.float32)
        for r in range(32):
            co = co0 + r
            if co < cout and nci > 0:
                base = (co * cin + ci0) * taps
                tile[r * ld:r * ld + nci * taps] = flat[base:base + nci * taps]
        for q in range(32 * taps):
            a, tp = q // taps, q % taps
            for lane in range(32):
                if co0 + a < cop and ci0 + lane < cip:
                    w[((co0 + a) * taps + tp) * cip + ci0 + lane] = tile[a * ld + lane * taps + tp]
                if ci0 + a < cip and co0 + lane < cop:
                    wt[((ci0 + a) * taps + tp) * cop + co0 + lane] = tile[lane * ld + a * taps + tp]
    return w, wt


@pytest.mark.parametrize("shape", [(64, 64, 9), (24, 40, 9), (6, 72, 9), (40, 27, 1), (72, 33, 9)])
def test_tile_pack_index_math(shape):
    cout, cin, taps = shape
    rng = np.random.RandomState(0)
    w32 = rng.randn(cout, cin, taps).astype(np.float32)
    w, wt = _replay(w32)
    cop, cip = _pad8(cout), _pad8(cin)
    ref_w = np.zeros((cop, taps, cip), np.float32)
    ref_w[:cout, :, :cin] = w32.transpose(0, 2, 1)
    ref_wt = np.zeros((cip, taps, cop), np.float32)
    ref_wt[:cin, :, :cout] = w32.transpose(1, 2, 0)
    assert np.array_equal(w, ref_w.reshape(-1))
    assert np.array_equal(wt, ref_wt.reshape(-1))
