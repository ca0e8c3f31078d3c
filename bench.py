#!/usr/bin/env python
"""Benchmark of the NPPNet hot path (BASELINE.json metric: train img/s @384^2 on 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A step = forward + Criterion_par + Criterion_pose + backward (+ gradient all-reduce) + Adam update of the derived
NPPNet (model_augment.Network, TRAIN.LAYERS=16, INIT_CHANNELS=64) on one batch of 32 synthetic LIP-shaped
384x384 images per GPU (BASELINE.json configs[1] / configs[2]).  Prints ONE JSON line on rank 0.
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

METRIC = "train img/s @384^2 (derived NPPNet, bf16, fwd+bwd+criteria+Adam)"
UNIT = "img/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (BASELINE configs[1]: 32)")
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=2)
    return ap.parse_args()


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


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_train_step_rate(args, steps, warmup, batch):
    """The reference's CPU path for the same step, via the oracle port (plain PyTorch fp32 restatement of
    model_augment.Network + core/criterion.py; /root/reference itself cannot travel to the GPU box).
    Returns (img/s, cores, seconds per step)."""
    import torch
    from npp_b200 import engine
    from npp_b200.models.model_augment import Network
    from oracle import nppnet_ref as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = Network(engine.make_cfg(layers=args.layers, init_channels=args.channels))  # parameters only (CPU)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    del net
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k and v.dim() > 0]
    lam_p = (2.3 * torch.ones(2)).requires_grad_(True)
    lam_q = (-2.5 * torch.ones(2)).requires_grad_(True)
    opt = torch.optim.Adam(params + [lam_p, lam_q], 0.0015)
    img, par, edge, g0, g1 = engine.synthetic_batch(batch, args.size, seed=1)
    w = torch.tensor(O.WEIGHTS_LIP)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        pose_l, par_l = O.network_forward(sd, img, layers=args.layers, training=True)
        loss = (O.criterion_par(par_l, [par, edge], lam_p, w).unsqueeze(0) +
                O.criterion_pose(pose_l, [g0, g1], lam_q).unsqueeze(0)).mean()
        loss.backward()
        opt.step()
        float(loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, cores, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    batch = args.cpu_batch if total <= 24 else 1
    rate, cores, sec = cpu_train_step_rate(args, args.steps, args.warmup, batch)
    sample = "oracle port (plain PyTorch fp32 on CPU) of the same train step, batch %d @%dx%d per step, %d threads" % (
        batch, args.size, args.size, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "derived NPPNet train step (fwd+criteria+bwd+Adam), %dx%d, CPU sample batch %d" % (
            args.size, args.size, batch), "layers": args.layers, "init_channels": args.channels},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    if args.impl == "reference":
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

    from npp_b200 import _lib, build, engine, distributed
    from npp_b200 import functional as F_
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    from npp_b200.models.model_augment import Network
    if rank == 0 and not os.path.exists(_lib.LIB_PATH):
        build.build_lib()
    if world > 1:
        dist.barrier()
    _lib.lib()
    F_.set_compute_dtype(torch.bfloat16)

    torch.manual_seed(0)
    model = Network(engine.make_cfg(layers=args.layers, init_channels=args.channels)).to(dev).train()
    cpose = Criterion_pose(out_len=2, use_target_weight=False).to(dev)
    cpar = Criterion_par(out_len=2).to(dev)
    opt = engine.build_optimizer(model, cpose, cpar)
    if world > 1:
        distributed.enable_sync_bn(True)
    step = engine.TrainStep(model, cpose, cpar, opt, args.batch, args.size, use_graph=not args.no_graph,
                            world_size=world)
    host = engine.synthetic_batch(args.batch, args.size, seed=1 + rank, pin=True)
    step.load(*host)
    torch.cuda.synchronize()

    # ---- eager warm-up, then the live measurement of the dominant kernel class: every dense-conv ABI call of one
    # step is traced (arguments kept alive) and re-issued back to back — first the tcgen05 implicit-GEMM fprop+dgrad
    # launches, then the wgrad launches — inside a CUDA graph, timed with CUDA events on the launch stream.  The
    # tensors of different convs are distinct and together far larger than L2, so the launches run cold like in the step.
    eager = engine.TrainStep(model, cpose, cpar, opt, args.batch, args.size, use_graph=False, world_size=world)
    eager.images, eager.par_lab, eager.edge_lab = step.images, step.par_lab, step.edge_lab
    eager.pose_gt, eager.pose_aux_gt = step.pose_gt, step.pose_aux_gt
    torch.cuda.reset_peak_memory_stats()
    eager.run()
    torch.cuda.synchronize()
    peak_mem = torch.cuda.max_memory_allocated()
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

    prof = {"conv_gemm(fprop+dgrad)": time_class(("npp_conv2d_fwd", "npp_conv2d_dgrad")),
            "conv_wgrad": time_class(("npp_conv2d_wgrad", "npp_conv2d_wgrad_ws"))}
    # every other kernel class of the step (all HBM-bound), same method: the class's launches of one step re-issued
    # back to back inside a CUDA graph; algorithmic bytes = what the call's NHWC views span (each view once; no credit
    # for halos, re-reads, fp32 coefficient vectors or workspaces).  Not replayed: the optimizer update and the
    # BatchNorm finalize (they would advance parameters / running statistics), NCCL.
    skip = {"npp_conv2d_fwd", "npp_conv2d_dgrad", "npp_conv2d_wgrad", "npp_conv2d_wgrad_ws", "npp_adam_step", "npp_bn_finalize"}
    for name in sorted(set(t[0] for t in trace) - skip):
        prof[name] = time_class((name,))
    del trace
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

    # ---- end to end through the public API: H2D of the batch from pinned host memory, step, D2H of the loss
    losses = []

    def e2e_step_plain():
        step.load(*host, non_blocking=True)
        step.run()
        losses.append(float(step.loss.item()))

    def e2e_step_prefetch():
        # public API with input prefetch: run() consumes the batch staged on the device and the upload of the NEXT
        # batch (pinned host -> device, side stream) overlaps the step; one full upload + one loss read per step
        step.run()
        step.prefetch(*host)
        losses.append(float(step.loss.item()))

    e2e_step, e2e_mode = e2e_step_prefetch, "next batch's H2D overlaps the running step (TrainStep.prefetch)"
    try:
        step.prefetch(*host)
        for _ in range(2):
            e2e_step()
    except Exception as exc:   # fall back to the serial upload
        print("prefetch path failed (%r): serial H2D" % (exc,), file=sys.stderr)
        step._staged = False
        e2e_step, e2e_mode = e2e_step_plain, "serial: H2D, step, D2H"
        for _ in range(2):
            e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop()

    if rank == 0:
        gbatch = args.batch * world
        value = gbatch * args.steps / (ms_total * 1e-3)
        e2e = gbatch * args.steps / (ms_e2e * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json, sustained)" if peaks else "fallback (B200_PROFILING.md)"
        gemm = prof["conv_gemm(fprop+dgrad)"]
        # DRAM traffic of the dominant kernel class from the committed ncu capture of the same kernels (one eager step
        # under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, tools/gpu_final.sh -> tools/ncu_traffic.py):
        # bytes per launch, averaged over the class like `achieved`
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = tj["conv_gemm"]["dram_bytes_per_launch"]
        except Exception:
            pass
        achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
        kernels = {}
        for name, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            if d["ms"] < 0.05:
                continue
            k = {"launches_per_step": d["calls"], "ms_per_step": round(d["ms"], 3)}
            if d["flops"]:
                k["tflops"] = round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1)
                k["tensor_frac"] = round(k["tflops"] / tf_peak, 3)
            elif d["bytes"]:
                k["gbs"] = round(d["bytes"] / (d["ms"] * 1e-3) / 1e9, 1)
                k["hbm_frac"] = round(k["gbs"] / hbm_peak, 3)
            kernels[name] = k
        bw = [d for n_, d in prof.items() if not d["flops"] and d["bytes"]]
        bw_ms, bw_bytes = sum(d["ms"] for d in bw), sum(d["bytes"] for d in bw)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "derived NPPNet (model_augment.Network, genotypes ENCODER/DECODER/INTER/FUSION) "
                                   "train step fwd+Criterion_par+Criterion_pose+bwd+Adam, batch %d/GPU @%dx%d, "
                                   "synthetic LIP-shaped images/labels, random init" % (args.batch, args.size, args.size),
                       "layers": args.layers, "init_channels": args.channels, "global_batch": gbatch,
                       "parallelism": "dp%d" % world, "cuda_graph": not args.no_graph,
                       "l2": "no flush: per-step working set (%.1f GB peak activations) >> 126 MB L2" % (peak_mem / 1e9)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": step.input_bytes(), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "loss_first": losses[0], "loss_last": losses[-1],
                    "mode": e2e_mode},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf_peak if tf_peak else None, "traffic": traffic,
                         "kernel": "conv_gemm2_kernel<BN> + conv3_kernel<BN> (tcgen05 implicit GEMM: fprop + dgrad, "
                                   "%d launches/step)" % gemm["calls"],
                         "how": "algorithmic FLOPs (2*N*Ho*Wo*Cout*Cin*kh*kw) of every dense-conv fprop and dgrad launch of "
                                "one step / CUDA-event time of those launches replayed back to back on the launch stream",
                         "avg_launch_us": 1e3 * gemm["ms"] / max(1, gemm["calls"]), "peak_source": peak_src,
                         "share_of_step": gemm["ms"] / (ms_total / args.steps)},
            "kernels": kernels,
            "kernel_ms_total": round(sum(d["ms"] for d in prof.values()), 3),
            "hbm_class": {"ms_per_step": round(bw_ms, 3), "gbs": round(bw_bytes / (bw_ms * 1e-3) / 1e9, 1) if bw_ms else None,
                          "frac": round(bw_bytes / (bw_ms * 1e-3) / 1e9 / hbm_peak, 3) if bw_ms else None,
                          "what": "all bandwidth-bound kernel classes of the step together (algorithmic bytes / time)"},
            "hbm_peak_gbs": hbm_peak,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                rate, cores, sec = cpu_train_step_rate(args, 2, 1, args.cpu_batch)
                line["cpu_baseline"] = {
                    "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "oracle port (plain PyTorch fp32) of the same train step on the host CPU, batch %d @%dx%d, "
                              "1 warm-up + 2 timed steps (%.1f s/step)" % (args.cpu_batch, args.size, args.size, sec)}
            except Exception as e:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing the NCCL communicator down: the captured CUDA graphs still reference it, and
        # destroy_process_group() / interpreter teardown was observed to block for minutes on that.  Every rank has
        # finished its timed work (the timing all-reduce above is the last collective); exit code 0 for torchrun.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
