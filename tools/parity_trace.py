"""Per-stage forward error trace of the bf16 product path against the oracle (fp64 truth; oracle with bf16-rounded
storage as the yardstick) at any configuration — locates WHERE a parity gap opens instead of judging the outputs only.

  python tools/parity_trace.py [--layers 16 --channels 64 --batch 8 --size 384] [--out gpurun_out/parity_trace.txt]

Stage names are shared by npp_b200.models.model_augment.Network._tr and oracle.nppnet_ref._tr.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.double(), b.double().to(a.device)
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.models.model_augment import Network
    from oracle import nppnet_ref as O
    F_.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(0)
    net = Network(engine.make_cfg(layers=a.layers, init_channels=a.channels))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.randn(a.batch, 3, a.size, a.size, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]

    mine = {}
    net = net.cuda().train()
    net._trace = lambda n, t: mine.__setitem__(n, t.float().cpu())
    with torch.no_grad():
        pl, par = net(x.cuda())
    for n, t in zip(names, [t for p in pl + par for t in p]):
        mine[n] = t.float().cpu()
    del net, pl, par
    torch.cuda.empty_cache()

    def oracle(dt, storage):
        store = {}
        O.set_trace(store)
        O.set_storage_dtype(storage)
        try:
            with torch.no_grad():
                s = {k: (v.cuda().to(dt) if v.is_floating_point() else v.cuda()) for k, v in sd.items()}
                pl, par = O.network_forward(s, x.cuda().to(dt), layers=a.layers, training=True)
        finally:
            O.set_trace(None)
            O.set_storage_dtype(None)
        for n, t in zip(names, [t for p in pl + par for t in p]):
            store[n] = t
        return {k: v.cpu() for k, v in store.items()}

    truth = oracle(torch.float64, None)
    torch.cuda.empty_cache()
    yard = oracle(torch.float32, torch.bfloat16)
    lines = ["%-28s %10s %10s" % ("stage (L=%d C=%d B=%d %d^2)" % (a.layers, a.channels, a.batch, a.size), "ours", "yardstick")]
    for k in truth:
        if k in mine:
            lines.append("%-28s %10.5f %10.5f" % (k, rel(mine[k], truth[k]), rel(yard[k], truth[k])))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
