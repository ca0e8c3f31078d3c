"""CPU replay of the work assignment of the two-group conv epilogue (csrc/conv_tcgen05.cu, EPI2 paths of
conv_gemm2_kernel and conv3_kernel): every 128-row x 64-column unit of every tile must be drained exactly once, by the
group the accumulator-release protocol expects, and every tile must collect exactly the number of mbarrier arrivals its
"TMEM empty" barrier was initialised with (256 threads = both groups, or 128 when the groups alternate tiles)."""
import pytest


def gemm2_epi2(bn, cout, nb, tiles):
    """Mirrors the EPI2 branch of conv_gemm2_kernel for one persistent CTA that owns `tiles` pixel tiles."""
    slabs = bn // 64 if bn >= 64 else 1
    nslab = min((cout - nb * bn + 63) // 64, slabs)
    gs = slabs // 2 if slabs >= 2 else 1
    alt = slabs < 2
    units, arrivals = [], {}
    for gi in (0, 1):
        acc = gi if alt else 0
        it = gi if alt else 0
        while it < tiles:
            released = False
            for s2 in range(gs):
                slab = 0 if alt else gi + 2 * s2
                if slab < nslab:
                    last = alt or slab + 2 >= nslab
                    units.append((it, slab, gi, acc))
                    if last:
                        arrivals[it] = arrivals.get(it, 0) + 128
                        released = True
            if not released:
                arrivals[it] = arrivals.get(it, 0) + 128
            if alt:
                it += 2
            else:
                it += 1
                acc ^= 1
    expect = 128 if bn <= 64 else 256
    return units, arrivals, expect, nslab, alt


@pytest.mark.parametrize("bn,cout", [(32, 32), (32, 24), (64, 64), (64, 40), (128, 128), (128, 72), (128, 64),
                                     (256, 256), (256, 192), (256, 136), (256, 512), (256, 384), (128, 320)])
@pytest.mark.parametrize("tiles", [1, 2, 5, 16])
def test_gemm2_units_and_arrivals(bn, cout, tiles):
    n_blocks = (cout + bn - 1) // bn
    for nb in range(n_blocks):
        units, arrivals, expect, nslab, alt = gemm2_epi2(bn, cout, nb, tiles)
        want = {(it, slab) for it in range(tiles) for slab in range(nslab)}
        got = [(it, slab) for it, slab, _, _ in units]
        assert sorted(got) == sorted(want)                       # every unit once
        assert all(arrivals[it] == expect for it in range(tiles)), (arrivals, expect)
        for it, slab, gi, acc in units:
            assert acc == (it & 1)                               # the stage the issuer of tile `it` writes
            if alt:
                assert gi == (it & 1)
            else:
                assert gi == (slab & 1)


def conv3_epi2(bn, cout, tiles):
    slabs = bn // 64
    nslab = min((cout + 63) // 64, slabs)
    units, arrivals = [], {}
    for gi in (0, 1):
        for it in range(tiles):
            if slabs == 1:
                units.append((it, gi, 0, gi))                    # (tile, pixel half m, slab, group)
                arrivals[it] = arrivals.get(it, 0) + 128
            else:
                slab = gi
                if slab < nslab:
                    for m in (0, 1):
                        units.append((it, m, slab, gi))
                    arrivals[it] = arrivals.get(it, 0) + 128     # released with the m == 1 unit
                else:
                    arrivals[it] = arrivals.get(it, 0) + 128
    return units, arrivals, nslab


@pytest.mark.parametrize("bn,cout", [(64, 64), (64, 40), (128, 128), (128, 72), (128, 64)])
def test_conv3_units_and_arrivals(bn, cout):
    tiles = 7
    units, arrivals, nslab = conv3_epi2(bn, cout, tiles)
    want = {(it, m, slab) for it in range(tiles) for m in (0, 1) for slab in range(nslab)}
    assert sorted((it, m, slab) for it, m, slab, _ in units) == sorted(want)
    assert all(arrivals[it] == 256 for it in range(tiles))
