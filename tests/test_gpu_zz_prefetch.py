"""GPU: TrainStep.prefetch — the input path bench.py's `e2e` figure goes through (host -> device copy of the next
batch on a side stream while the current step runs)."""
import pytest
import torch

from test_gpu_engine import _make

pytestmark = pytest.mark.gpu


def test_prefetch_feeds_the_same_batches_as_load(lib_built):
    """TrainStep.prefetch (H2D of the next batch on a side stream, consumed by the next run()) must feed the step
    exactly what load() would: with the learning rate at 0 the graph replays see the same parameters, so the losses
    of the two input paths agree batch by batch (5e-3: BatchNorm sums are fp32 atomics)."""
    from npp_b200 import engine
    batches = [engine.synthetic_batch(2, 128, seed=20 + i, pin=True) for i in range(4)]
    model, step = _make(3, use_graph=True)
    for g in step.opt.param_groups:
        g["lr"] = 0.0
    step.load(*batches[0])
    step.prepare()
    ref = []
    for b in batches:
        step.load(*b)
        ref.append(float(step.run()))
    got = []
    step.prefetch(*batches[0])
    for i in range(len(batches)):
        step.run()                                   # consumes the staged batch i
        if i + 1 < len(batches):
            step.prefetch(*batches[i + 1])           # uploads batch i+1 while step i runs
        got.append(float(step.loss.item()))
    torch.cuda.synchronize()
    for a, b in zip(ref, got):
        assert abs(a - b) <= 5e-3 * abs(a), (ref, got)
    assert len(set(round(v, 4) for v in ref)) > 1   # the batches really differ
