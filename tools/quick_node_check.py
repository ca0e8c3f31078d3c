"""Ten-second device check of the node backward kernels (both compile-time variants: with / without the second
gradient inputs) against plain torch autograd on the GPU.  Prints PASS/FAIL lines; exit code 1 on failure.

    python tools/quick_node_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from npp_b200 import functional as F_  # noqa: E402
from npp_b200.nn import BatchNorm2d  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
F_.set_compute_dtype(torch.bfloat16)
n, c, h, w = 4, 32, 24, 24
fails = 0
for want_cat in (False, True):
    a32 = torch.randn(n, c, h, w, device=dev).bfloat16().float()
    b32 = torch.randn(n, c, h, w, device=dev).bfloat16().float()
    bn_a, bn_b = BatchNorm2d(c).to(dev).train(), BatchNorm2d(c).to(dev).train()
    with torch.no_grad():
        for bn in (bn_a, bn_b):
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
    gs = [torch.randn(n, c, h, w, device=dev).bfloat16().float() for _ in range(4)]
    # ---- ours
    a = F_.to_internal(a32.clone().requires_grad_(True))
    b = F_.to_internal(b32.clone().requires_grad_(True))
    pa, pb = F_.Pending(a, None, bn_a), F_.Pending(b, None, bn_b)
    if want_cat:
        raw, rel, raw_c, rel_c = F_.node(pa, pb, want_raw=True, want_relu=True, want_cat=True)
        outs = [raw, rel, raw_c, rel_c]
    else:
        raw, rel = F_.node(pa, pb, want_raw=True, want_relu=True)
        outs = [raw, rel]
    loss = sum((o.float() * g).sum() for o, g in zip(outs, gs))
    da, db = torch.autograd.grad(loss, [a, b])
    # ---- torch reference (fp32 math on the same bf16-rounded inputs)
    ar, br = a32.clone().requires_grad_(True), b32.clone().requires_grad_(True)

    def bn_ref(x, bn):
        return torch.nn.functional.batch_norm(x, None, None, bn.weight.detach(), bn.bias.detach(), True, 0.1, bn.eps)

    y = bn_ref(ar, bn_a) + bn_ref(br, bn_b)
    refs = [y, torch.relu(y)] + ([y, torch.relu(y)] if want_cat else [])
    lref = sum((o * g).sum() for o, g in zip(refs, gs))
    dar, dbr = torch.autograd.grad(lref, [ar, br])
    for name, got, ref in (("raw", raw, y), ("da", da, dar), ("db", db, dbr)):
        err = ((got.detach().float() - ref.detach()).norm() / ref.detach().norm()).item()
        ok = err < 2e-2
        fails += not ok
        print("%s want_cat=%s %-3s rel err %.3e" % ("PASS" if ok else "FAIL", want_cat, name, err), flush=True)
torch.cuda.synchronize()
sys.exit(1 if fails else 0)
