"""Runs the dominant dense-conv shapes of the bench workload (fprop + dgrad + wgrad through the C ABI) for ncu:

  ncu --set full --clock-control none --import-source on -k regex:conv_ -c 24 -o gpurun_out/conv_full \
      python tools/profile_conv.py
Also prints CUDA-event timings per shape and direction (TFLOP/s) when run without a profiler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from npp_b200 import functional as F_  # noqa: E402

F_.set_compute_dtype(torch.bfloat16)
B = int(os.environ.get("NPP_PROF_BATCH", "32"))
SHAPES = [  # (cin, cout, k, H, stride)  — SURVEY.md §2a
    (128, 128, 3, 96, 1), (256, 256, 3, 48, 1), (1024, 512, 1, 96, 1), (32, 32, 3, 96, 1), (512, 128, 1, 96, 1),
    (128, 128, 3, 24, 1), (256, 256, 3, 12, 1), (64, 64, 3, 48, 1), (128, 128, 1, 96, 1), (64, 128, 3, 192, 2),
]
if os.environ.get("NPP_PROF_SHAPES"):   # e.g. "128,128,3,96,1;32,32,3,96,1"
    SHAPES = [tuple(int(v) for v in t.split(",")) for t in os.environ["NPP_PROF_SHAPES"].split(";")]
reps = int(os.environ.get("NPP_PROF_REPS", "1"))
for cin, cout, k, h, s in SHAPES:
    x = F_.to_internal(torch.randn(B, cin, h, h, device="cuda"), torch.bfloat16).detach().requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, device="cuda") * 0.05).requires_grad_(True)
    ho = (h + 2 * (k // 2) - k) // s + 1
    gy = F_.to_internal(torch.randn(B, cout, ho, ho, device="cuda"), torch.bfloat16)
    flops = 2.0 * B * ho * ho * cout * cin * k * k
    if not os.environ.get("NPP_PROF_NOWARM"):
        y, _ = F_.conv2d(x, w, None, s, k // 2, 1, want_stats=True)   # warm-up (tensor maps, attributes)
        y.backward(gy)
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tf = tb = 0.0
    for _ in range(reps):
        x.grad = w.grad = None
        ev[0].record()
        y, _ = F_.conv2d(x, w, None, s, k // 2, 1, want_stats=True)
        ev[1].record()
        ev[2].record()
        y.backward(gy)
        ev[3].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1])
        tb += ev[2].elapsed_time(ev[3])
    print("conv %4d->%4d k%d s%d @%3d^2 B=%d: fwd(+pack) %.1f us  %.0f TFLOP/s | dgrad+wgrad %.1f us %.0f TFLOP/s" % (
        cin, cout, k, s, h, B, 1e3 * tf / reps, flops / (tf / reps * 1e-3) / 1e12, 1e3 * tb / reps,
        2 * flops / (tb / reps * 1e-3) / 1e12), flush=True)
