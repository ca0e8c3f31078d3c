"""Debug: per-parameter gradient difference between eager launches and CUDA-graph replay of the same step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from npp_b200 import engine
from npp_b200 import functional as F_
from npp_b200.core.criterion import Criterion_par, Criterion_pose
from npp_b200.models.model_augment import Network


def make(seed, use_graph):
    F_.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(seed)
    model = Network(engine.make_cfg(layers=8, init_channels=16)).cuda().train()
    cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2, min_kept=2000).cuda()
    opt = engine.build_optimizer(model, cpose, cpar)
    for g in opt.param_groups:
        g["lr"] = 0.0
    return model, engine.TrainStep(model, cpose, cpar, opt, 2, 128, use_graph=use_graph, warmup=1)


b0 = engine.synthetic_batch(2, 128, seed=10)
b1 = engine.synthetic_batch(2, 128, seed=11)
m1, e1 = make(0, False)
m2, e2 = make(0, False)
m3, g3 = make(0, True)
g3.load(*b0); g3.prepare()
for st in (e1, e2, g3):
    st.load(*b1)
l1, l2, l3 = float(e1.run()), float(e2.run()), float(g3.run())
l3b = float(g3.run())
torch.cuda.synchronize()
print("losses eager/eager/graph/graph2", l1, l2, l3, l3b)
def rel(a, b): return ((a - b).norm() / (a.norm() + 1e-30)).item()
print("flat eager-vs-eager", rel(e1.flat_grads, e2.flat_grads), "eager-vs-graph", rel(e1.flat_grads, g3.flat_grads))
rows = []
for (k, p), q, r in zip(m1.named_parameters(), m2.parameters(), m3.parameters()):
    if p.grad is None: continue
    rows.append((rel(p.grad, r.grad), rel(p.grad, q.grad), p.grad.norm().item(), k))
rows.sort(reverse=True)
for r in rows[:25]:
    print("%.3e (eager2 %.3e) |g|=%.3e %s" % r)
