"""Device parity of the tcgen05 implicit-GEMM convolution kernels (fprop / BatchNorm statistics / dgrad / wgrad)
against the CUDA-core direct convolution of the same library, through the C ABI, via the standalone binary
tests/csrc/test_conv.cu (one process per case: a deadlocked mbarrier pipeline must not take the suite down).

Kernel selections exercised (switches documented in npp_b200/csrc/conv_tcgen05.cu):
  default  eight-warp epilogue + the 256-pixel halo-sharing 3x3 kernel where the launch is large enough (weights
           resident in shared memory for Cin <= 64, weight ring otherwise; two MMA issuer warps alternating tiles);
  forced   the 3x3 kernel on every 3x3 / stride-1 case it can tile, however small or ragged;
  ring     forced, but weights always through the ring (the path the wide layers take, here on narrow ones too);
  oneprod  ring with a single TMA producer warp feeding both rings;
  legacy   the first-generation kernels only.
The n32 "bench shape" cases give every persistent CTA several tiles (both issuer warps wrap their rings).
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "csrc", "_bin", "test_conv")

MODES = {
    "default": {},
    "forced": {"NPP_CONV3_MIN_TILES": "1", "NPP_CONV3_PAD_PCT": "400"},
    "ring": {"NPP_CONV3_MIN_TILES": "1", "NPP_CONV3_PAD_PCT": "400", "NPP_CONV3_WRES": "0"},
    "oneprod": {"NPP_CONV3_MIN_TILES": "1", "NPP_CONV3_PAD_PCT": "400", "NPP_CONV3_WRES": "0", "NPP_CONV3_2PROD": "0"},
    "legacy": {"NPP_CONV_EPI8": "0", "NPP_CONV3": "0"},
}
# small cases only (the n32 bench shapes are timed by tools/, not here); indices into cases[] of test_conv.cu
SMALL = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 26, 27, 29, 30, 31]
FORCED = [2, 3, 4, 12, 13, 16, 26, 27, 29, 30, 31]   # 3x3 / stride 1 / Cout <= 128
LEGACY = [0, 3, 6, 8, 27]
RING = [2, 4, 12, 27, 29, 21]         # Cin <= 64: resident weights by default, the ring here
ONEPROD = [2, 3, 27]
MULTI_TILE = [17, 21, 23, 32, 33]     # n32 bench shapes: 2-16 tiles per persistent CTA


def _params():
    out = [("default", i) for i in SMALL + MULTI_TILE] + [("forced", i) for i in FORCED] + [("legacy", i) for i in LEGACY]
    out += [("ring", i) for i in RING] + [("oneprod", i) for i in ONEPROD]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("mode,case", _params())
def test_conv_case(mode, case, lib_built):
    if not os.path.exists(BIN):
        from npp_b200 import build
        build.build_test_binaries()
    env = dict(os.environ)
    env.update(MODES[mode])
    r = subprocess.run([BIN, str(case)], env=env, capture_output=True, text=True, timeout=90)
    assert r.returncode == 0, "test_conv case %d (%s) failed:\n%s\n%s" % (case, mode, r.stdout, r.stderr)
    assert "FAIL" not in r.stdout, r.stdout
