"""GPU: the tile-transposing multi-tensor weight pack (npp_pack_weights_tiles, NPP_PACK_TILES=1 — a round-2
candidate, NOT on the default path) against the per-tensor pack kernel the conv parity tests are built on
(npp_pack_weight): both packed layouts must be bit-identical, padding included.  Written after round 1's GPU budget
was spent, hence the non-strict xfail (a pass shows up as XPASS)."""
import struct

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="candidate kernel not yet run on a B200 (written after the round-1 GPU budget)",
                                strict=False)]

SHAPES = [(64, 64, 3), (128, 128, 3), (32, 32, 3), (24, 40, 3), (6, 384, 3), (512, 1024, 1), (20, 128, 1), (64, 27, 1)]


def test_tile_pack_matches_single_pack(lib_built):
    from npp_b200 import functional as F_
    from npp_b200._lib import call, fptr, i32, stream, NPP_BF16
    torch.manual_seed(4)
    rows, tt, tix, keep = [], [], [], []
    for cout, cin, k in SHAPES:
        w = torch.randn(cout, cin, k, k, device="cuda")
        cop, cip, taps = F_.pad8(cout), F_.pad8(cin), k * k
        n = cop * taps * cip
        ref_w, ref_wt = torch.empty(n, dtype=torch.bfloat16, device="cuda"), torch.empty(n, dtype=torch.bfloat16, device="cuda")
        call("npp_pack_weight", fptr(w), fptr(ref_w), fptr(ref_wt), i32(cout), i32(taps), i32(cin), i32(cop), i32(cip),
             i32(NPP_BF16), stream())
        got_w = torch.full((n,), 7.0, dtype=torch.bfloat16, device="cuda")
        got_wt = torch.full((n,), 7.0, dtype=torch.bfloat16, device="cuda")
        rows.append(struct.pack("<QQQiiiiii", w.data_ptr(), got_w.data_ptr(), got_wt.data_ptr(), cout, taps, cin, cop, cip, 0))
        nt = ((cop + 31) // 32) * ((cip + 31) // 32)
        tt += [len(rows) - 1] * nt
        tix += list(range(nt))
        keep.append((w, ref_w, ref_wt, got_w, got_wt))
    table = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).clone().cuda()
    tt_d, tix_d = torch.tensor(tt, dtype=torch.int32).cuda(), torch.tensor(tix, dtype=torch.int32).cuda()
    call("npp_pack_weights_tiles", fptr(table), i32(len(rows)), fptr(tt_d), fptr(tix_d), i32(len(tt)), stream())
    torch.cuda.synchronize()
    for (cout, cin, k), (w, ref_w, ref_wt, got_w, got_wt) in zip(SHAPES, keep):
        assert torch.equal(got_w.view(torch.int16), ref_w.view(torch.int16)), (cout, cin, k, "w")
        assert torch.equal(got_wt.view(torch.int16), ref_wt.view(torch.int16)), (cout, cin, k, "wt")
