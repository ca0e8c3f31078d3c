"""Per-(ABI call, shape) time table of one training step of the bench workload.

One eager step is traced at the C ABI (npp_b200/_lib.py trace_*), the recorded calls are grouped by entry point +
tensor shapes + integer arguments, and every group is re-issued back to back (all its calls, REPS times) between
CUDA events.  Prints groups sorted by total time with algorithmic TFLOP/s or GB/s.

  python tools/shape_table.py [--batch 32] [--top 60] > gpurun_out/shape_table.txt
"""
import argparse
import ctypes
import os
import sys
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from npp_b200 import _lib, engine  # noqa: E402
from npp_b200 import functional as F_  # noqa: E402
from npp_b200.core.criterion import Criterion_par, Criterion_pose  # noqa: E402
from npp_b200.models.model_augment import Network  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=384)
ap.add_argument("--layers", type=int, default=16)
ap.add_argument("--channels", type=int, default=64)
ap.add_argument("--top", type=int, default=70)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

F_.set_compute_dtype(torch.bfloat16)
torch.manual_seed(0)
model = Network(engine.make_cfg(layers=args.layers, init_channels=args.channels)).cuda().train()
cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2).cuda()
opt = engine.build_optimizer(model, cpose, cpar)
step = engine.TrainStep(model, cpose, cpar, opt, args.batch, args.size, use_graph=False)
step.load(*engine.synthetic_batch(args.batch, args.size, seed=1))
step.run()
torch.cuda.synchronize()
_lib.trace_begin()
step.run()
trace = _lib.trace_end()
torch.cuda.synchronize()

SKIP = {"npp_adam_step", "npp_bn_finalize"}


def key_of(entry):
    name, cargs, _, (pending, _keep) = entry
    shapes = tuple(tuple(t.shape) for t in pending if t.dim() == 4)
    ints = tuple(a.value for a in cargs if isinstance(a, ctypes.c_int))
    return (name, shapes, ints)


groups = OrderedDict()
for e in trace:
    if e[0] in SKIP:
        continue
    groups.setdefault(key_of(e), []).append(e)

cur = _lib.stream()
rows = []
for key, calls in groups.items():
    fn = getattr(_lib.lib(), key[0])
    for _ in range(1):
        for _, cargs, _, _ in calls:
            fn(*cargs[:-1], cur)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        for _, cargs, _, _ in calls:
            fn(*cargs[:-1], cur)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    fl = sum(c[2][0] for c in calls)
    by = sum(c[2][1] for c in calls)
    rows.append((ms, key, len(calls), fl, by))

rows.sort(key=lambda r: -r[0])
total = sum(r[0] for r in rows)
print("groups %d, calls %d, total %.2f ms (back-to-back, same buffers: small groups run L2-warm)" % (
    len(rows), sum(r[2] for r in rows), total))
print("%8s %6s %9s %9s  %s" % ("ms", "calls", "us/call", "rate", "call / shapes / ints"))
for ms, key, n, fl, by in rows[:args.top]:
    if fl:
        rate = "%7.1f TF" % (fl / (ms * 1e-3) / 1e12)
    elif by:
        rate = "%7.0f GB" % (by / (ms * 1e-3) / 1e9)
    else:
        rate = "        -"
    print("%8.3f %6d %9.1f %s  %s %s %s" % (ms, n, 1e3 * ms / n, rate, key[0], list(key[1]), list(key[2])))
