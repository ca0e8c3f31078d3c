"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF (read-only at /root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference ships no tests / golden vectors, so these are outputs of its own modules on
seeded inputs; they pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py).
Everything is stored as small compressed .npz files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refshim  # noqa: E402


def seeded_state(module, seed):
    """Deterministic, init-order-independent parameters: every tensor drawn from its own key-derived stream."""
    import zlib
    with torch.no_grad():
        for k, v in module.state_dict().items():
            g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(k.encode()))
            if k.endswith("num_batches_tracked"):
                continue
            if k.endswith("running_var"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif v.dim() == 1 and k.endswith("weight"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif v.dim() == 1:
                v.copy_(torch.randn(v.shape, generator=g) * 0.2)
            else:
                v.copy_(torch.randn(v.shape, generator=g) * (1.5 / v[0].numel() ** 0.5))


def np_state(module):
    return {"sd/" + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def golden_ops(ref):
    out = {}
    C, N, H, W = 16, 2, 12, 12
    cases = [("std_conv_3x3", 1), ("std_conv_3x3", 2), ("std_conv_1x1", 1), ("dil_conv_3x3_2", 1), ("dil_conv_3x3_4", 2),
             ("se_connect", 1), ("se_connect", 2), ("max_pool_3x3", 1), ("max_pool_3x3", 2), ("skip_connect", 2),
             ("poled_conv_x1", 1), ("sep_conv_3x3", 1), ("avg_pool_3x3", 1)]
    for i, (name, stride) in enumerate(cases):
        op = ref.operations.OPS[name](C, stride, True)
        seeded_state(op, 100 + i)
        op.train()
        g = torch.Generator().manual_seed(500 + i)
        x = torch.randn(N, C, H, W, generator=g).bfloat16().float().requires_grad_(True)
        tag = "%s_s%d" % (name, stride)
        for k, v in np_state(op).items():
            out["%s/%s" % (tag, k)] = v
        y = op(x)
        gy = torch.randn(y.shape, generator=g)
        (y * gy).sum().backward()
        out[tag + "/x"] = x.detach().numpy()
        out[tag + "/gy"] = gy.numpy()
        out[tag + "/y"] = y.detach().numpy()
        out[tag + "/dx"] = x.grad.numpy()
        for k, p in op.named_parameters():
            if p.grad is not None:
                out["%s/grad/%s" % (tag, k)] = p.grad.numpy()
        for k, b in op.named_buffers():
            if "running" in k:
                out["%s/after/%s" % (tag, k)] = b.detach().numpy()
    out["cases"] = np.array(["%s_s%d" % c for c in cases])
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **out)


def golden_network(ref):
    """Derived Network (model_augment.py) at the smallest shape every kernel supports: L=8, C=16, 2x3x128x128 (64x64 leaves 2x2 maps at the coarsest scale: BN over 8 samples is too ill-conditioned to pin anything)."""
    cfg = _refshim.cfg(layers=8, init_channels=16)
    torch.manual_seed(0)
    net = ref.model_augment.Network(cfg)
    net.train()
    # a checksum of the seeded init so the test can tell "init differs" from "forward differs"
    checksum = float(sum(v.double().sum() for k, v in net.state_dict().items() if v.is_floating_point()))
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 3, 128, 128, generator=g).bfloat16().float()
    pose_list, par_list = net(x)
    out = {"x": x.numpy(), "seed": np.array(0), "layers": np.array(8), "channels": np.array(16)}
    names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
    tensors = [t for pair in pose_list + par_list for t in pair]
    for n_, t in zip(names, tensors):
        out["out/" + n_] = t.detach().numpy()
    out["init_checksum"] = np.array([checksum])
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), **out)


def golden_loss(ref):
    torch.Tensor.cuda = lambda self, *a, **k: self  # criterion.py:192,197 hard-code .cuda()
    g = torch.Generator().manual_seed(9)
    B, H, W, LH, LW = 2, 12, 12, 48, 48
    par = [[(torch.randn(B, 20, H, W, generator=g) * 2).requires_grad_(True),
            (torch.randn(B, 2, H, W, generator=g)).requires_grad_(True)] for _ in range(2)]
    pose = [[torch.rand(B, 16, H, W, generator=g).requires_grad_(True),
             torch.rand(B, 16, H, W, generator=g).requires_grad_(True)] for _ in range(2)]
    lab = torch.randint(0, 20, (B, LH, LW), generator=g)
    lab[:, :3, :] = 255
    lab[:, :, -2:] = 255
    edge = (torch.rand(B, LH, LW, generator=g) < 0.07).long()
    edge[lab == 255] = 255
    gt = [torch.rand(B, 16, H, W, generator=g), torch.rand(B, 16, H, W, generator=g)]
    out = {"lab": lab.numpy(), "edge": edge.numpy(), "gt0": gt[0].numpy(), "gt1": gt[1].numpy()}
    for min_kept, tag in ((131072, "default"), (300, "kept300")):
        cp = ref.criterion.Criterion_par(out_len=2, ignore_index=255, thres=0.9, min_kept=min_kept)
        cq = ref.criterion.Criterion_pose(out_len=2, use_target_weight=False)
        for t in [p for pair in par + pose for p in pair]:
            t.grad = None
        lp = cp(par, [lab, edge])
        lq = cq(pose, gt)
        (lp + lq).backward()
        out[tag + "/loss_par"] = lp.detach().numpy()
        out[tag + "/loss_pose"] = lq.detach().numpy()
        out[tag + "/dlamda_par"] = cp.lamda.grad.numpy()
        out[tag + "/dlamda_pose"] = cq.lamda.grad.numpy()
        for i in range(2):
            out["%s/dpar%d" % (tag, i)] = par[i][0].grad.numpy().copy()
            out["%s/dedge%d" % (tag, i)] = par[i][1].grad.numpy().copy()
            out["%s/dpose%d" % (tag, i)] = pose[i][0].grad.numpy().copy()
            out["%s/dposeaux%d" % (tag, i)] = pose[i][1].grad.numpy().copy()
    for i in range(2):
        out["par%d" % i] = par[i][0].detach().numpy()
        out["edgelogit%d" % i] = par[i][1].detach().numpy()
        out["pose%d" % i] = pose[i][0].detach().numpy()
        out["poseaux%d" % i] = pose[i][1].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **out)


def golden_eval(ref):
    sys.modules.setdefault("matplotlib", type(sys)("matplotlib"))
    import importlib
    rng = np.random.RandomState(3)
    out = {}
    # confusion matrix (utils/utils.py:192-218) — import the function without the rest of utils' heavy imports
    src = open(os.path.join(_refshim.REF_ROOT, "utils", "utils.py")).read()
    start = src.index("def get_confusion_matrix")
    end = src.index("def adjust_learning_rate")
    ns = {"np": np, "torch": torch}
    exec(compile(src[start:end], "ref_utils_get_confusion_matrix", "exec"), ns)
    logits = rng.randn(2, 20, 24, 24).astype(np.float32)
    logits[0, 3, :4, :4] = logits[0, 7, :4, :4] = 9.0  # exact ties: first maximum must win
    label = rng.randint(0, 20, size=(2, 26, 26)).astype(np.int64)
    label[:, :2, :] = 255
    cm = ns["get_confusion_matrix"](torch.from_numpy(label), torch.from_numpy(logits), (2, 20, 24, 24), 20, 255)
    out.update({"cm/logits": logits, "cm/label": label, "cm/matrix": cm})
    # accuracy (core/evaluate.py:68-99)
    hm = rng.rand(5, 16, 16, 12).astype(np.float32)
    gt = rng.rand(5, 16, 16, 12).astype(np.float32)
    gt[:, -1] = 0           # a joint that is never annotated (the reference's __main__ recipe, evaluate.py:141-143)
    gt[1, 2] = 0
    hm[2, 5] = -1.0         # max <= 0 -> prediction zeroed
    for b in range(5):      # make some predictions land close to the target
        for j in range(0, 16, 3):
            hm[b, j] = gt[b, j] + 0.01 * rng.rand(16, 12)
    acc, avg_acc, cnt, pred = ref.evaluate.accuracy(hm, gt)
    out.update({"acc/hm": hm, "acc/gt": gt, "acc/acc": acc, "acc/avg": np.array(avg_acc), "acc/cnt": np.array(cnt),
                "acc/pred": pred})
    # PCKh (utils/calc_pckh.py) on the reference's shipped fixture rows
    pckh = importlib.import_module("utils.calc_pckh")
    gt_xy, _ = pckh.read_data(os.path.join(_refshim.REF_ROOT, "prepare_files", "pose_csv", "pose_gt.csv"), True)
    gt_xy = gt_xy[:400]
    pred_xy = gt_xy + rng.randn(*gt_xy.shape) * 12.0
    pred_xy[gt_xy < 0] = 1  # read_data(pred, False) maps negatives to 1
    dist = pckh.get_norm_dist(pred_xy, gt_xy, pckh.get_head_size(gt_xy))
    pck = pckh.compute_pck(dist, np.array([0.5]))
    out.update({"pckh/gt": gt_xy, "pckh/pred": pred_xy, "pckh/pck": pck})
    # flip-test merge (core/function.py:927-939): executed exactly as written there
    import torch.nn.functional as F
    p = torch.from_numpy(rng.randn(2, 20, 12, 12).astype(np.float32))
    fp = torch.from_numpy(rng.randn(2, 20, 12, 12).astype(np.float32))
    pred_par = F.interpolate(input=p, size=(48, 48), mode="bilinear")
    flip_pred_par = F.interpolate(input=fp, size=(48, 48), mode="bilinear")
    tmp = flip_pred_par
    flip_pred_par[:, 14, :, :] = tmp[:, 15, :, :]
    flip_pred_par[:, 15, :, :] = tmp[:, 14, :, :]
    flip_pred_par[:, 16, :, :] = tmp[:, 17, :, :]
    flip_pred_par[:, 17, :, :] = tmp[:, 16, :, :]
    flip_pred_par[:, 18, :, :] = tmp[:, 19, :, :]
    flip_pred_par[:, 19, :, :] = tmp[:, 18, :, :]
    flip_pred_par = flip_pred_par.flip(3)
    merged = 0.5 * (pred_par + flip_pred_par)
    out.update({"tta/pred": p.numpy(), "tta/flip": fp.numpy(), "tta/merged": merged.numpy()})
    np.savez_compressed(os.path.join(HERE, "eval_golden.npz"), **out)


SEARCH_GRAD_KEYS = ["stem0.0.weight", "cells1.3._ops.0.net.1.weight", "_ops1.0._ops.0.0.net.1.weight",
                    "_ops2.1.extra_conv.weight", "up_ops1.7._ops.4.0.net.2.weight", "pose_net.0._ops.5._ops.5.net.1.weight",
                    "par_net.2.preprocess1.net.1.weight", "par_head.1.4.bias"]


def golden_search(ref):
    """Search supernet (model_search_interact.py) at L=8, C=16, 2x3x128x128 with random architecture tensors:
    outputs, d(alphas/betas) and a few weight gradients of  sum_i <out_i, r_i>  (r_i seeded)."""
    import importlib
    rs = importlib.import_module("models.model_search_interact")
    cfg = _refshim.cfg(layers=8, init_channels=16)
    torch.manual_seed(0)
    net = rs.Network(cfg)
    net.train()
    g = torch.Generator().manual_seed(123)
    arch_names = ["alphas1", "alphas2", "alphas3", "alphas4", "alphas_pose", "alphas_par", "betas1", "betas2",
                  "betas3", "betas4", "betas_pose", "betas_par"]
    out = {"layers": np.array(8), "channels": np.array(16), "seed": np.array(0)}
    for n_, pa in zip(arch_names, net.arch_parameters()):
        pa.data.copy_(torch.randn(pa.shape, generator=g) * 0.5)
        out["arch/" + n_] = pa.detach().numpy().copy()
    x = torch.randn(2, 3, 128, 128, generator=g).bfloat16().float()
    out["x"] = x.numpy()
    pose_list, par_list = net(x)
    names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
    tensors = [t for pair in pose_list + par_list for t in pair]
    gr = torch.Generator().manual_seed(321)
    loss = 0
    for n_, t in zip(names, tensors):
        out["out/" + n_] = t.detach().numpy()
        loss = loss + (t * torch.randn(t.shape, generator=gr)).sum()
    loss = loss + 3.0 * net.loss_entropy()
    loss.backward()
    out["loss"] = np.array([float(loss)])
    out["entropy"] = np.array([float(net.loss_entropy())])
    for n_, pa in zip(arch_names, net.arch_parameters()):
        out["grad/" + n_] = pa.grad.numpy().copy()
    sd = dict(net.named_parameters())
    for k in SEARCH_GRAD_KEYS:
        out["wgrad/" + k] = sd[k].grad.numpy().copy()
    gi, gf = net.genotype()
    out["genotype"] = np.array([repr((gi, [list(gf.pose), list(gf.par)]))])
    np.savez_compressed(os.path.join(HERE, "search_golden.npz"), **out)


if __name__ == "__main__":
    assert _refshim.have_reference(), "needs /root/reference"
    ref = _refshim.import_reference()
    torch.set_num_threads(8)
    golden_ops(ref)
    golden_network(ref)
    golden_eval(ref)
    golden_loss(ref)
    golden_search(ref)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
