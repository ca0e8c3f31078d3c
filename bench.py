#!/usr/bin/env python
"""Benchmark of the NPPNet hot path (BASELINE.json metric: train img/s @384^2 on 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W                 # our arm, BASELINE configs[1] / [2] (torchrun for N > 1)
  python bench.py --workload search   ...                       # configs[3]: supernet fwd+bwd with alphas, B=16/GPU @384^2
  python bench.py --workload infer512 ...                       # configs[4]: inference @512^2 + GPU mIoU / PCK evaluation
  python bench.py --impl reference     ...                      # the reference's CPU path (oracle port) on the host cores
  python bench.py --impl reference-gpu ...                      # the same oracle port on the B200 through stock torch
                                                                #   (cuDNN / ATen): fp32 and bf16 autocast + channels_last

train    : a step = forward + Criterion_par + Criterion_pose + backward (+ gradient all-reduce) + Adam update of the
           derived NPPNet (model_augment.Network, TRAIN.LAYERS=16, INIT_CHANNELS=64), batch 32 per GPU.
search   : a step = the weight step of the search supernet (model_search_interact.Network, SEARCH.LAYERS=16,
           INIT_CHANNELS=32): forward + criteria + backward (weight AND architecture gradients) + Adam on the
           weights + Adam(0.5, 0.999) on the alphas / betas; --bilevel times the two-batch `train_with_alpha` schedule.
infer512 : a step = two eval-mode forwards (image + mirror) of the derived NPPNet with 7 classes / 14 joints at 512^2,
           flip-merge, confusion histogram and PCK counters on the device (core/function_ppp.py:869-964).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRICS = {
    "train": "train img/s @384^2 (derived NPPNet, bf16, fwd+bwd+criteria+Adam)",
    "search": "search img/s @384^2 (supernet, bf16, fwd+bwd with alphas+criteria+Adam x2)",
    "infer512": "eval img/s @512^2 (derived NPPNet 7 cls / 14 joints, bf16, 2 forwards + flip merge + GPU mIoU/PCK)",
}
UNIT = "img/s"
DEFAULTS = {"train": (32, 384, 64), "search": (16, 384, 32), "infer512": (16, 512, 64)}   # batch, size, init channels


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="train", choices=sorted(METRICS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--channels", type=int, default=None)
    ap.add_argument("--bilevel", action="store_true", help="search: time the two-batch train_with_alpha schedule")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true", help="skip the per-kernel-class replay measurement")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--full-kernels", action="store_true", help="list every kernel class instead of the top 12")
    a = ap.parse_args()
    b, s, c = DEFAULTS[a.workload]
    a.batch = a.batch or b
    a.size = a.size or s
    a.channels = a.channels or c
    return a


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- oracle arms
def _oracle_setup(args, device):
    """Oracle-port state for the workload: (state_dict, forward fn, criteria lamdas, class weights, batch maker)."""
    import torch
    from npp_b200 import engine
    from oracle import nppnet_ref as O
    torch.manual_seed(0)
    if args.workload == "search":
        from npp_b200.models.model_search_interact import Network
        net = Network(engine.make_cfg(layers=args.layers, init_channels=args.channels))
        fwd = lambda sd, x, training=True: O.search_forward(sd, x, layers=args.layers, training=training)
        nc, nj = 20, 16
    else:
        from npp_b200.models.model_augment import Network
        nc, nj = (7, 14) if args.workload == "infer512" else (20, 16)
        net = Network(engine.make_cfg(num_classes=nc, num_joints=nj, layers=args.layers, init_channels=args.channels))
        fwd = lambda sd, x, training=True: O.network_forward(sd, x, layers=args.layers, training=training)
    sd = {k: v.detach().clone().to(device) for k, v in net.state_dict().items()}     # parameters only (built on CPU)
    del net
    return sd, fwd, nc, nj


def oracle_step_rate(args, steps, warmup, batch, device="cpu", autocast=False):
    """The reference's arithmetic for the same step through the oracle port (plain PyTorch restatement of
    model_augment / model_search_interact + core/criterion.py + core/evaluate.py; /root/reference itself cannot travel
    to the GPU box).  device='cpu': the host cores (cpu_baseline / --impl reference); device='cuda': stock torch on
    the GPU = cuDNN / ATen (--impl reference-gpu), fp32 or bf16 autocast + channels_last.
    Returns (img/s, seconds per step)."""
    import contextlib
    import torch
    from npp_b200 import engine
    from oracle import eval_ref as E
    from oracle import nppnet_ref as O
    sd, fwd, nc, nj = _oracle_setup(args, device)
    cuda = device != "cpu"
    if cuda and autocast:
        sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if (cuda and autocast) else contextlib.nullcontext
    train = args.workload != "infer512"
    if train:
        params = [v.requires_grad_(True) for k, v in sd.items()
                  if v.is_floating_point() and "running" not in k and v.dim() > 0]
        lam_p = (2.3 * torch.ones(2, device=device)).requires_grad_(True)
        lam_q = (-2.5 * torch.ones(2, device=device)).requires_grad_(True)
        arch = [v for k, v in sd.items() if k.startswith(("alphas", "betas"))]
        arch_ids = set(id(v) for v in arch)
        opt = torch.optim.Adam([p for p in params if id(p) not in arch_ids] + [lam_p, lam_q], 0.0015)
        a_opt = torch.optim.Adam(arch, lr=0.001, betas=(0.5, 0.999), weight_decay=0.001) if arch else None
        w = torch.tensor(O.WEIGHTS_LIP, device=device)
        img, par, edge, g0, g1 = [t.to(device) for t in engine.synthetic_batch(batch, args.size, seed=1)]
    else:
        g = torch.Generator().manual_seed(1)
        img = torch.randn(batch, 3, args.size, args.size, generator=g).to(device)
        par = torch.randint(0, nc, (batch, args.size, args.size), generator=g)
        gt = torch.rand(batch, nj, args.size // 4, args.size // 4, generator=g)
    if cuda and autocast:
        img = img.contiguous(memory_format=torch.channels_last)
    times = []
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if train:
            opt.zero_grad(set_to_none=True)
            if a_opt is not None:
                a_opt.zero_grad(set_to_none=True)
            with ctx():
                pose_l, par_l = fwd(sd, img)
            pose_l = [[t.float() for t in p] for p in pose_l]
            par_l = [[t.float() for t in p] for p in par_l]
            loss = (O.criterion_par(par_l, [par, edge], lam_p, w).unsqueeze(0) +
                    O.criterion_pose(pose_l, [g0, g1], lam_q).unsqueeze(0)).mean()
            loss.backward()
            opt.step()
            if a_opt is not None:
                a_opt.step()
            float(loss.detach())
        else:
            with torch.no_grad(), ctx():
                pose_l, par_l = fwd(sd, img, training=False)
                fpose_l, fpar_l = fwd(sd, img.flip(3), training=False)
            # the reference's evaluation runs on the host (function_ppp.py:923-960): D2H of the logits, numpy
            merged = E.tta_merge(par_l[-1][0].float(), fpar_l[-1][0].float(), (args.size, args.size), swap_lr=False)
            E.confusion_matrix(par.numpy(), merged.cpu().numpy(), (batch, nc, args.size, args.size), nc, 255)
            hm = E.flip_average_pascal(pose_l[-1][0].float().cpu().numpy(), fpose_l[-1][0].float().cpu().numpy())
            E.pck_counts(hm, gt.numpy())
        if cuda:
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec


def _workload_text(args, batch=None):
    b = batch if batch is not None else args.batch
    if args.workload == "train":
        return ("derived NPPNet (model_augment.Network, genotypes ENCODER/DECODER/INTER/FUSION) train step "
                "fwd+Criterion_par+Criterion_pose+bwd+Adam, batch %d/GPU @%dx%d, synthetic LIP-shaped images/labels, "
                "random init" % (b, args.size, args.size))
    if args.workload == "search":
        return ("search supernet (model_search_interact.Network: 164 MixedOps, alphas/betas) %s fwd+criteria+bwd (weight + "
                "architecture gradients)+Adam x2, batch %d/GPU @%dx%d, synthetic LIP-shaped data, random init" % (
                    "bilevel train_with_alpha step (2 batches)" if args.bilevel else "weight step", b, args.size, args.size))
    return ("derived NPPNet 7 classes / 14 joints, eval mode: forward(image) + forward(mirror) + flip merge + confusion "
            "histogram + PCK counters, batch %d/GPU @%dx%d, synthetic pascal-shaped data, random init" % (b, args.size, args.size))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = getattr(args, "workload", "train")
    if args.impl == "reference-gpu":
        return run_reference_gpu_arm(args)
    cores = os.cpu_count() or 1
    import torch
    torch.set_num_threads(cores)
    batch = args.cpu_batch                      # ONE sample size for every CPU leg (in-line cpu_baseline uses it too)
    rate, sec = oracle_step_rate(args, args.steps, args.warmup, batch, "cpu")
    sample = "oracle port (plain PyTorch fp32 on CPU) of the same step, batch %d @%dx%d per step, %d threads" % (
        batch, args.size, args.size, cores)
    line = {
        "impl": "reference", "metric": METRICS[workload], "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload_text(args, batch) + " [CPU sample batch %d]" % batch, "layers": args.layers,
                   "init_channels": args.channels},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_reference_gpu(args, steps, warmup):
    """Stock torch on the same GPU: {mode: {img/s, ms, batch}}.  Falls back to half the batch on OOM."""
    import torch
    out = {}
    for mode, autocast in (("fp32", False), ("bf16_autocast_channels_last", True)):
        batch = args.batch
        while batch >= 1:
            try:
                torch.cuda.empty_cache()
                rate, sec = oracle_step_rate(args, steps, warmup, batch, "cuda", autocast)
                out[mode] = {"value": rate, "unit": UNIT, "ms_per_step": sec * 1e3, "batch": batch}
                break
            except torch.cuda.OutOfMemoryError:
                batch //= 2
        else:
            out[mode] = {"value": None, "unit": UNIT, "note": "out of memory at every batch size"}
    out["what"] = ("oracle port of the same step (identical arithmetic to the reference modules) run by stock PyTorch %s on "
                   "this GPU: cuDNN / ATen kernels, eager launches, torch defaults (cudnn.allow_tf32=%s, matmul.allow_tf32=%s)"
                   % (torch.__version__, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32))
    return out


def run_reference_gpu_arm(args):
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    res = measure_reference_gpu(args, args.steps, max(args.warmup, 2))
    best = max((v for v in res.values() if isinstance(v, dict) and v.get("value")), key=lambda v: v["value"])
    line = {"impl": "reference-gpu", "metric": METRICS[args.workload], "value": best["value"], "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 2), "ms_per_step": best["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if best is res.get("bf16_autocast_channels_last") else "f32",
            "data": "synthetic", "config": {"workload": _workload_text(args, best["batch"]), "layers": args.layers,
                                            "init_channels": args.channels},
            "gpu_reference": res, "gpu_launches": 0,
            "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def build_step(args, dev, world):
    """(step object with load/prepare/run/input_bytes, eager twin sharing its inputs, host batch tuple)."""
    import torch
    from npp_b200 import distributed, engine
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    rank = int(os.environ.get("RANK", "0"))
    torch.manual_seed(0)
    if args.workload == "infer512":
        from npp_b200.models.model_augment import Network
        model = Network(engine.make_cfg(num_classes=7, num_joints=14, layers=args.layers,
                                        init_channels=args.channels)).to(dev)
        step = engine.EvalStep(model, args.batch, args.size, use_graph=not args.no_graph)
        eager = engine.EvalStep(model, args.batch, args.size, use_graph=False)
        g = torch.Generator().manual_seed(1 + rank)
        host = [torch.randn(args.batch, 3, args.size, args.size, generator=g).pin_memory(),
                torch.randint(0, 7, (args.batch, args.size, args.size), generator=g).pin_memory(),
                torch.rand(args.batch, 14, args.size // 4, args.size // 4, generator=g).pin_memory()]
        eager.images, eager.par_lab, eager.pose_gt = step.images, step.par_lab, step.pose_gt
        return step, eager, host
    cpose = Criterion_pose(out_len=2, use_target_weight=False).to(dev)
    cpar = Criterion_par(out_len=2).to(dev)
    if world > 1:
        distributed.enable_sync_bn(True)
    if args.workload == "search":
        from npp_b200.models.model_search_interact import Network
        model = Network(engine.make_cfg(layers=args.layers, init_channels=args.channels)).to(dev).train()
        w_opt, a_opt = engine.build_search_optimizers(model, cpose, cpar)
        mk = lambda graph: engine.SearchStep(model, cpose, cpar, w_opt, a_opt, args.batch, args.size, use_graph=graph,
                                             world_size=world, bilevel=args.bilevel)
    else:
        from npp_b200.models.model_augment import Network
        model = Network(engine.make_cfg(layers=args.layers, init_channels=args.channels)).to(dev).train()
        opt = engine.build_optimizer(model, cpose, cpar)
        mk = lambda graph: engine.TrainStep(model, cpose, cpar, opt, args.batch, args.size, use_graph=graph,
                                            world_size=world)
    step, eager = mk(not args.no_graph), mk(False)
    eager.images, eager.par_lab, eager.edge_lab = step.images, step.par_lab, step.edge_lab
    eager.pose_gt, eager.pose_aux_gt = step.pose_gt, step.pose_aux_gt
    if args.workload == "search" and args.bilevel:
        eager._in2 = step._in2
    host = engine.synthetic_batch(args.batch, args.size, seed=1 + rank, pin=True)
    return step, eager, host


CONV_FWD = ("npp_conv2d_fwd", "npp_conv2d_dgrad")
CONV_WGRAD = ("npp_conv2d_wgrad", "npp_conv2d_wgrad_ws")
# never replayed: they advance parameters / running statistics, or talk to the other ranks
NO_REPLAY = {"npp_adam_step", "npp_bn_finalize", "npp_peer_allreduce", "npp_node_fwd_bn"}


def main():
    args = parse()
    if args.impl != "ours":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from npp_b200 import _lib, build, distributed
    from npp_b200 import functional as F_
    if rank == 0 and not os.path.exists(_lib.LIB_PATH):
        build.build_lib()
    if world > 1:
        dist.barrier()
    _lib.lib()
    F_.set_compute_dtype(torch.bfloat16)

    step, eager, host = build_step(args, dev, world)
    step.load(*host)
    if args.workload == "search" and args.bilevel:
        step.load2(*host)
    torch.cuda.synchronize()

    # ---- eager warm-up, then the live measurement of the kernel classes: every ABI call of one eager step is traced
    # (arguments kept alive) and each class's launches are re-issued back to back inside a CUDA graph, timed with CUDA
    # events on the launch stream.  The tensors of different launches are distinct and together far larger than L2, so
    # the launches run cold like in the step.
    torch.cuda.reset_peak_memory_stats()
    eager.run()
    torch.cuda.synchronize()
    peak_mem = torch.cuda.max_memory_allocated()
    prof = {}
    if not args.no_kernel_table:
        _lib.trace_begin()
        eager.run()
        trace = _lib.trace_end()
        torch.cuda.synchronize()

        def time_class(names, reps=3):
            calls = [t for t in trace if t[0] in names]
            if not calls:
                return {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0}
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                _lib.replay(trace, names)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                _lib.replay(trace, names)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return {"calls": len(calls), "ms": e0.elapsed_time(e1) / reps, "flops": sum(t[2][0] for t in calls),
                    "bytes": sum(t[2][1] for t in calls)}

        prof = {"conv_gemm(fprop+dgrad)": time_class(CONV_FWD), "conv_wgrad": time_class(CONV_WGRAD)}
        # every other kernel class of the step (all HBM-bound), same method; algorithmic bytes = what the call's NHWC
        # views span (each view once; no credit for halos, re-reads, fp32 coefficient vectors or workspaces)
        for name in sorted(set(t[0] for t in trace) - set(CONV_FWD) - set(CONV_WGRAD) - NO_REPLAY):
            prof[name] = time_class((name,))
        # node_fwd_bn also finalizes BatchNorm (running statistics): replayed with the statistics restored afterwards
        if any(t[0] == "npp_node_fwd_bn" for t in trace):
            saved = [b.detach().clone() for b in eager.model.buffers()]
            prof["npp_node_fwd_bn"] = time_class(("npp_node_fwd_bn",))
            with torch.no_grad():
                for b, s_ in zip(eager.model.buffers(), saved):
                    b.copy_(s_)
        del trace
    if world > 1 and distributed.peer_comm() is not None:
        distributed.peer_comm().check()
    torch.cuda.empty_cache()

    # ---- capture + warm-up
    step.prepare()
    for _ in range(max(args.warmup, 3)):
        step.run()
    torch.cuda.synchronize()

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    c0 = _lib.launch_count()
    ms_total = timed(step.run, args.steps)
    launches = (_lib.launch_count() - c0) if args.no_graph else step.launches_per_step * args.steps

    # ---- end to end through the public API: H2D of the batch from pinned host memory, step, D2H of the result
    results = []
    train_like = args.workload != "infer512"

    def read_result():
        if train_like:
            results.append(float(step.loss.item()))          # 4 B: the loss scalar (core/function.py:114)
        else:
            results.append(int(step.hist.sum().item()))      # forces the counters; the full read is in d2h_bytes

    def e2e_step_plain():
        step.load(*host, non_blocking=True)
        step.run()
        read_result()

    def e2e_step_prefetch():
        # public API with input prefetch: run() consumes the batch staged on the device and the upload of the NEXT
        # batch (pinned host -> device, side stream) overlaps the step; one full upload + one result read per step
        step.run()
        step.prefetch(*host)
        read_result()

    e2e_step, e2e_mode = e2e_step_plain, "serial: H2D, step, D2H"
    if hasattr(step, "prefetch") and not (args.workload == "search" and args.bilevel):
        try:
            step.prefetch(*host)
            e2e_step_prefetch()
            e2e_step, e2e_mode = e2e_step_prefetch, "next batch's H2D overlaps the running step (TrainStep.prefetch)"
        except Exception as exc:   # fall back to the serial upload
            print("prefetch path failed (%r): serial H2D" % (exc,), file=sys.stderr)
            step._staged = False
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop()
    peer_seq = None
    if world > 1 and distributed.peer_comm() is not None:
        peer_seq = distributed.peer_comm().check()

    if rank == 0:
        gbatch = args.batch * world
        per_step_imgs = gbatch * (2 if (args.workload == "search" and args.bilevel) else 1)
        value = per_step_imgs * args.steps / (ms_total * 1e-3)
        e2e = per_step_imgs * args.steps / (ms_e2e * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json, sustained)" if peaks else "fallback (B200_PROFILING.md)"
        d2h = 4 if train_like else (step.hist.numel() + 2 * step.hit.numel()) * 8
        line = {
            "metric": METRICS[args.workload], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": _workload_text(args), "layers": args.layers, "init_channels": args.channels,
                       "global_batch": gbatch, "parallelism": "dp%d" % world, "cuda_graph": not args.no_graph,
                       "l2": "no flush: per-step working set (%.1f GB peak activations) >> 126 MB L2" % (peak_mem / 1e9),
                       "syncbn": (None if world == 1 else ("nvlink peer-memory one-shot all-reduce (csrc/peer.cu), %d "
                                                           "exchanges so far" % peer_seq if peer_seq else "nccl all-reduce"))},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": step.input_bytes(), "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "result_first": results[0], "result_last": results[-1],
                    "mode": e2e_mode},
            "gpu_launches": int(launches),
        }
        if prof:
            gemm, wg = prof["conv_gemm(fprop+dgrad)"], prof["conv_wgrad"]
            # DRAM traffic of the dominant kernel class: bytes per launch from the committed ncu per-launch capture of
            # the SAME kernels (profiles/roofline_traffic.json names the build it was taken from)
            traffic, traffic_src = None, None
            if args.workload == "train":      # the capture is one eager step of THIS workload
                try:
                    tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
                    traffic = tj["conv_gemm"]["dram_bytes_per_launch"]
                    traffic_src = tj.get("source") or tj["conv_gemm"].get("source")
                except Exception:
                    pass
            else:
                traffic_src = "not captured for this workload (profiles/roofline_traffic.json is the train step)"
            achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
            wg_tf = wg["flops"] / (wg["ms"] * 1e-3) / 1e12 if wg["ms"] > 0 else 0.0
            kernels = {}
            for name, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
                if d["ms"] < 0.05:
                    continue
                k = {"n": d["calls"], "ms": round(d["ms"], 2)}
                if d["flops"]:
                    k["tflops"] = round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1)
                    k["frac"] = round(k["tflops"] / tf_peak, 3)
                elif d["bytes"]:
                    k["gbs"] = round(d["bytes"] / (d["ms"] * 1e-3) / 1e9)
                    k["frac"] = round(k["gbs"] / hbm_peak, 3)
                kernels[name.replace("npp_", "")] = k
            if not args.full_kernels:
                kernels = dict(list(kernels.items())[:12])      # the record keeps what explains the step; the rest on request
            bw = [d for n_, d in prof.items() if not d["flops"] and d["bytes"]]
            bw_ms, bw_bytes = sum(d["ms"] for d in bw), sum(d["bytes"] for d in bw)
            line["roofline"] = {
                "bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": achieved / tf_peak if tf_peak else None, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "conv_gemm2_kernel<BN> + conv3_kernel<BN> (tcgen05 implicit GEMM: fprop + dgrad, %d launches/step)"
                          % gemm["calls"],
                "how": "algorithmic FLOPs (2*N*Ho*Wo*Cout*Cin*kh*kw) of every dense-conv fprop and dgrad launch of one "
                       "step / CUDA-event time of those launches replayed back to back on the launch stream",
                "avg_launch_us": 1e3 * gemm["ms"] / max(1, gemm["calls"]), "peak_source": peak_src,
                "share_of_step": gemm["ms"] / (ms_total / args.steps),
                "wgrad": {"achieved": wg_tf, "frac": wg_tf / tf_peak if tf_peak else None, "launches": wg["calls"],
                          "ms_per_step": round(wg["ms"], 3),
                          "kernel": "conv_wgrad3_kernel<BN> / conv_wgrad_kernel<BN> + wgrad_reduce_kernel (split-K fold)"},
                "hbm_class": {"ms_per_step": round(bw_ms, 3),
                              "achieved": round(bw_bytes / (bw_ms * 1e-3) / 1e9, 1) if bw_ms else None, "peak": hbm_peak,
                              "unit": "GB/s", "frac": round(bw_bytes / (bw_ms * 1e-3) / 1e9 / hbm_peak, 3) if bw_ms else None,
                              "what": "all bandwidth-bound kernel classes of the step together (algorithmic bytes / time)"},
            }
            # nested under `roofline` so the per-class table survives record keepers that drop unknown top-level keys
            line["roofline"]["classes"] = kernels
            line["roofline"]["classes_ms_total"] = round(sum(d["ms"] for d in prof.values()), 2)
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                rate, sec = oracle_step_rate(args, 2, 1, args.cpu_batch, "cpu")
                line["cpu_baseline"] = {
                    "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "oracle port (plain PyTorch fp32) of the same step on the host CPU, batch %d @%dx%d, 1 warm-up "
                              "+ 2 timed steps (%.1f s/step)" % (args.cpu_batch, args.size, args.size, sec)}
            except Exception as e:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
    # ---- stock torch on the same GPU (the honest bar): after our memory is released
    if world == 1 and not args.no_gpu_reference:
        step.close() if hasattr(step, "close") else None
        del step, eager
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            line["gpu_reference"] = measure_reference_gpu(args, 3, 2)
            best = max((v["value"] for v in line["gpu_reference"].values() if isinstance(v, dict) and v.get("value")),
                       default=None)
            if best:
                line["gpu_reference"]["speedup_vs_best"] = round(value / best, 2)
            if isinstance(line.get("cpu_baseline"), dict):      # a compact copy beside the other baseline
                line["cpu_baseline"]["gpu_reference"] = {k: (round(v["value"], 1) if isinstance(v, dict) and v.get("value")
                                                             else None) for k, v in line["gpu_reference"].items()
                                                         if k in ("fp32", "bf16_autocast_channels_last")}
                line["cpu_baseline"]["gpu_reference"]["ours_over_best"] = line["gpu_reference"].get("speedup_vs_best")
        except Exception as e:
            line["gpu_reference"] = {"failed": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # orderly teardown: captured graphs reference the NCCL communicator and the peer-memory buffers, so they go
        # first; a watchdog turns a stuck destroy_process_group() into a clean exit instead of a hung job
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(60.0, lambda: os._exit(0)).start()
        torch.cuda.synchronize()
        step.close()
        del eager
        distributed.enable_sync_bn(None)
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    main()
