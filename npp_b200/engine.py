"""Training / inference step drivers: the caller side of the hot path (core/function.py:57-120 `train`,
augment_lip_sync.py:186-212 model/criterion/optimizer setup) arranged for B200.

The reference issues ~4 000 kernel launches per step from Python; here the whole step — forward, both
criteria, backward, gradient all-reduce and optimizer update — is captured once into a CUDA graph and
replayed, so the host only enqueues one graph launch plus the H2D copies of the next batch.
"""
import gc
import types

import torch
import torch.nn as nn

from . import _lib
from . import functional as F_


def make_cfg(num_classes=20, num_joints=16, layers=16, init_channels=64, refine_layers=1):
    """The 8 config fields Network reads (core/config.py defaults + experiments/lip/384_384.yaml)."""
    ns = types.SimpleNamespace
    return ns(DATASET=ns(NUM_CLASSES=num_classes, NUM_JOINTS=num_joints),
              TRAIN=ns(LAYERS=layers, INIT_CHANNELS=init_channels),
              SEARCH=ns(LAYERS=layers, INIT_CHANNELS=init_channels),
              MODEL=ns(DECONV_WITH_BIAS=False, HEAD="PSP", REFINE_LAYERS=refine_layers))


def build_optimizer(model, criterion_pose, criterion_par, lr=0.0015, fused=True):
    """Adam with the reference's parameter groups (augment_lip_sync.py:193-212): backbone at 0.2*LR,
    the rest at LR, the criteria's uncertainty weights at 1e-4.  fused=True uses npp_b200.optim.FusedAdam
    (one launch per step), fused=False torch.optim.Adam(capturable=True)."""
    def backbone(n):
        return n.startswith("cells1.") or n.startswith("cells2") or n.startswith("stem")
    groups = [
        {"params": [p for n, p in model.named_parameters() if backbone(n) and p.requires_grad], "lr": 0.2 * lr},
        {"params": [p for n, p in model.named_parameters() if not backbone(n) and p.requires_grad]},
    ]
    if fused:
        from .optim import FusedAdam
        opt = FusedAdam(groups, lr)
    else:
        opt = torch.optim.Adam(groups, lr, capturable=True, foreach=True)
    opt.add_param_group({"params": list(criterion_pose.parameters()), "lr": 0.0001})
    opt.add_param_group({"params": list(criterion_par.parameters()), "lr": 0.0001})
    return opt


def synthetic_batch(batch, size=384, num_classes=20, num_joints=16, seed=1, device="cpu", pin=False):
    """LIP-shaped synthetic batch (dataset/data_loader.py:285-304, core/function.py:73-84):
    images fp32 [B,3,S,S]; parsing labels int64 [B,S,S] in {0..C-1} with a 255 border band; edge labels
    {0,1,255} with P(1)~0.05; two fp32 [B,J,S/4,S/4] heat-map targets built from Gaussian blobs."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, 3, size, size, generator=g)
    par = torch.randint(0, num_classes, (batch, size, size), generator=g)
    band = max(2, size // 48)
    par[:, :band, :] = 255
    par[:, -band:, :] = 255
    par[:, :, :band] = 255
    par[:, :, -band:] = 255
    edge = (torch.rand(batch, size, size, generator=g) < 0.05).long()
    edge[par == 255] = 255
    hs = size // 4
    ys = torch.arange(hs).view(1, 1, hs, 1).float()
    xs = torch.arange(hs).view(1, 1, 1, hs).float()
    cy = torch.rand(batch, num_joints, 1, 1, generator=g) * hs
    cx = torch.rand(batch, num_joints, 1, 1, generator=g) * hs
    pose = []
    for sigma in (7.0 / 4.0 * 1.0, 14.0 / 4.0 * 1.0):
        pose.append(torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * sigma * sigma)))
    out = [img, par, edge, pose[0], pose[1]]
    if pin:
        out = [t.pin_memory() for t in out]
    if device != "cpu":
        out = [t.to(device) for t in out]
    return out


class TrainStep:
    """One data-parallel training step of the derived NPPNet:
        pose, par = model(images); loss = mean(criterion_par(par, [par_lab, edge_lab]) + criterion_pose(pose, [gt, gt_aux]))
        loss.backward(); (all-reduce grads); optimizer.step()          (core/function.py:87-107)
    With use_graph=True the step is captured into a CUDA graph after `warmup` eager steps."""

    def __init__(self, model, criterion_pose, criterion_par, optimizer, batch, size=384, use_graph=True,
                 world_size=1, warmup=3):
        self.model, self.cpose, self.cpar, self.opt = model, criterion_pose, criterion_par, optimizer
        self.opts = [optimizer]      # every optimizer whose state a step touches (SearchStep adds the alpha optimizer)
        self.world_size = world_size
        self.use_graph = use_graph
        dev = next(model.parameters()).device
        nj = model._num_joints
        hs = size // 4
        self.images = torch.zeros(batch, 3, size, size, device=dev)
        self.par_lab = torch.zeros(batch, size, size, dtype=torch.int64, device=dev)
        self.edge_lab = torch.zeros(batch, size, size, dtype=torch.int64, device=dev)
        self.pose_gt = torch.zeros(batch, nj, hs, hs, device=dev)
        self.pose_aux_gt = torch.zeros(batch, nj, hs, hs, device=dev)
        self.loss = torch.zeros((), device=dev)
        self.graph = None
        self._warmup = max(1, warmup) if use_graph else warmup
        # persistent flat gradients: static addresses for the graph, one memset per step, all-reduce without copies
        self.flat_grads = optimizer.use_flat_grads() if hasattr(optimizer, "use_flat_grads") else None
        self.launches_per_step = None
        self.packer = F_.WeightPacker(model) if dev.type == "cuda" else None   # all conv weights: one launch per step
        self._stage = None       # prefetch(): staging copies of the inputs, filled on a side stream
        self._staged = False
        if world_size > 1:
            # SyncBN here scales the local pixel count by the group size (one 2C-float message per BatchNorm instead of
            # torch's (mean, invstd, count) gather), which is only right when every rank holds the same batch
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                sizes = [None] * dist.get_world_size()
                dist.all_gather_object(sizes, (int(batch), int(size)))
                if len(set(sizes)) != 1:
                    raise RuntimeError("TrainStep needs the same per-rank batch on every rank (SyncBN statistics assume "
                                       "equal sample counts; use drop_last), got %s" % (sizes,))

    # ---- pieces ------------------------------------------------------------------------------------
    def load(self, images, par_lab, edge_lab, pose_gt, pose_aux_gt, non_blocking=True):
        """Copies a batch (host pinned or device) into the static input buffers."""
        self.images.copy_(images, non_blocking=non_blocking)
        self.par_lab.copy_(par_lab, non_blocking=non_blocking)
        self.edge_lab.copy_(edge_lab, non_blocking=non_blocking)
        self.pose_gt.copy_(pose_gt, non_blocking=non_blocking)
        self.pose_aux_gt.copy_(pose_aux_gt, non_blocking=non_blocking)

    def _inputs(self):
        return (self.images, self.par_lab, self.edge_lab, self.pose_gt, self.pose_aux_gt)

    def prefetch(self, images, par_lab, edge_lab, pose_gt, pose_aux_gt):
        """Uploads the NEXT batch (pinned host memory) into staging buffers on a side stream, so the host->device
        copy overlaps the step that is running; the next run() moves it into the static inputs with device-to-device
        copies (the DataLoader prefetch of augment_lip_sync.py:173-184, moved onto the device side)."""
        if self._stage is None:
            self._stage = [torch.empty_like(t) for t in self._inputs()]
            self._copy_stream = torch.cuda.Stream()
            self._ev_up, self._ev_used = torch.cuda.Event(), torch.cuda.Event()
            self._ev_used.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ev_used)   # the previous staged batch has been consumed
            for s, h in zip(self._stage, (images, par_lab, edge_lab, pose_gt, pose_aux_gt)):
                s.copy_(h, non_blocking=True)
            self._ev_up.record(self._copy_stream)
        self._staged = True

    def _consume_staged(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ev_up)
        for d, s in zip(self._inputs(), self._stage):
            d.copy_(s, non_blocking=True)
        self._ev_used.record(cur)
        self._staged = False

    def input_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.images, self.par_lab, self.edge_lab, self.pose_gt,
                                                           self.pose_aux_gt))

    def _allreduce_grads(self, opt=None):
        """DistributedDataParallel's gradient averaging (augment_lip_sync.py:207-208) on the flat buffer."""
        import torch.distributed as dist
        if opt is None:
            opt, flat = getattr(self, "opt", None), self.flat_grads
        else:
            flat = self.flat_grads if opt is getattr(self, "opt", None) else getattr(opt, "flat_grads", None)
        if flat is not None:
            if flat.is_cuda and dist.get_backend() == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)      # the 1 / world scaling rides on the collective
            else:
                dist.all_reduce(flat)
                flat.div_(self.world_size)
            return
        params = [p for g in opt.param_groups for p in g["params"] if p.grad is not None]
        grads = [p.grad for p in params]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat)
        flat.div_(self.world_size)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)

    def _step_body(self):
        self.opt.zero_grad(set_to_none=True)
        F_._arena.begin(self.images.device)             # every zero-initialised accumulator of the step: one memset
        F_._state["defer_bn_counters"] = counters = []  # BatchNorm num_batches_tracked += 1: one launch, not 432
        try:
            if self.packer is not None:
                self.packer.pack()
            pose, par = self.model(self.images)
            F_._state["defer_bn_counters"] = None
            if counters:
                torch._foreach_add_(counters, 1)
            loss_par = self.cpar(par, [self.par_lab, self.edge_lab]).unsqueeze(0)
            loss_pose = self.cpose(pose, [self.pose_gt, self.pose_aux_gt]).unsqueeze(0)
            loss = (loss_par + loss_pose).mean()
            loss.backward()
        finally:
            F_._state["defer_bn_counters"] = None
            F_._arena.end()
            F_.WeightPacker.release()
        if self.world_size > 1:
            self._allreduce_grads()
        self.opt.step()
        self.loss.copy_(loss.detach())

    # ---- warm-up without side effects ----------------------------------------------------------------------
    def _trainable_state(self):
        """Every tensor a training step mutates, keyed by what it IS rather than where it lives (the optimizer may
        re-seat its step counters on the first step after a load_state_dict): parameters (model + criteria), buffers
        (BatchNorm running statistics, num_batches_tracked) and the optimizers' moments / step counters."""
        out = {}
        for m in (self.model, self.cpose, self.cpar):
            for p in m.parameters():
                out[("param", id(p))] = p
            for b in m.buffers():
                out[("buffer", id(b))] = b
        for oi, opt in enumerate(self.opts):
            for p, st in opt.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        out[("opt", oi, id(p), k)] = v
        return out

    # ---- public ------------------------------------------------------------------------------------------
    def prepare(self, keep_state=True):
        """Eager warm-up steps (cuTensorMap driver entry point, kernel attributes, optimizer state allocation, arena
        sizing) and graph capture.  keep_state=True: the warm-up steps run on whatever is in the input buffers, so
        their effect on the weights, BatchNorm running statistics and Adam moments / step counters is UNDONE afterwards
        (values restored in place, addresses unchanged): training starts from the state prepare() found, one update
        per batch like the reference (core/function.py:105-107)."""
        if self.graph is not None or not self.use_graph:
            if not self.use_graph and self.launches_per_step is None:
                snap = {k: t.detach().clone() for k, t in self._trainable_state().items()} if keep_state else None
                c0 = _lib.launch_count()
                self._step_body()
                self.launches_per_step = _lib.launch_count() - c0
                if snap is not None:
                    self._restore(snap)
            return
        gc.collect()  # autograd graphs of earlier steps (AccumulateGrad nodes bound to another stream) must be gone
        snap = {k: t.detach().clone() for k, t in self._trainable_state().items()} if keep_state else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self._warmup):
                self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        F_._arena.reserve(self.images.device)   # sized by the warm-up step; cannot grow during capture
        self.graph = torch.cuda.CUDAGraph()
        c0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._step_body()
        self.launches_per_step = _lib.launch_count() - c0
        if snap is not None:
            self._restore(snap)

    def _restore(self, snap):
        with torch.no_grad():
            for k, t in self._trainable_state().items():
                s = snap.get(k)
                if s is not None:
                    t.copy_(s.to(dtype=t.dtype, device=t.device))   # e.g. a loaded float / CPU `step` re-seated as int64
                else:
                    t.zero_()     # optimizer state created by the warm-up (exp_avg, exp_avg_sq, step): back to initial

    def run(self):
        """Runs one step on whatever is in the static input buffers (or on the batch staged by prefetch()); returns
        the (device) loss scalar."""
        if self._staged:
            self._consume_staged()
        if self.use_graph:
            if self.graph is None:
                self.prepare()
            for opt in self.opts:
                if hasattr(opt, "sync_hyperparameters"):
                    opt.sync_hyperparameters()   # lr schedulers: the graph reads lr / weight_decay from the device table
            self.graph.replay()
        else:
            self._step_body()
        return self.loss

    def close(self):
        """Drops the captured graph (it references the NCCL communicator and the peer-memory buffers): call before
        torch.distributed.destroy_process_group() / distributed.enable_sync_bn(None)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
        gc.collect()


def build_search_optimizers(model, criterion_pose, criterion_par, w_lr=0.0015, alpha_lr=0.001, fused=True):
    """The two optimizers of search_lip_sync.py:273-279: Adam over every non-architecture parameter (+ the criteria's
    uncertainty weights at 1e-4) and Adam(lr=APLHA_LR, betas=(0.5, 0.999), weight_decay=1e-3) over arch_parameters()."""
    arch = set(map(id, model.arch_parameters()))
    weights = [p for p in model.parameters() if id(p) not in arch and p.requires_grad]
    if fused:
        from .optim import FusedAdam
        w_opt = FusedAdam(weights, w_lr)
        a_opt = FusedAdam(list(model.arch_parameters()), lr=alpha_lr, betas=(0.5, 0.999), weight_decay=0.001)
    else:
        w_opt = torch.optim.Adam(weights, w_lr, capturable=True, foreach=True)
        a_opt = torch.optim.Adam(list(model.arch_parameters()), lr=alpha_lr, betas=(0.5, 0.999), weight_decay=0.001,
                                 capturable=True, foreach=True)
    w_opt.add_param_group({"params": list(criterion_pose.parameters()), "lr": 0.0001})
    w_opt.add_param_group({"params": list(criterion_par.parameters()), "lr": 0.0001})
    return w_opt, a_opt


class SearchStep(TrainStep):
    """One step of the supernet search (core/function.py:485-621 `train_with_alpha`):

        weight step on batch 1:  loss1 = mean(criterion_par + criterion_pose);  optimizer.zero_grad / backward / step
        alpha  step on batch 2:  loss2 = 2 * mean(criterion_par + criterion_pose [+ 2 * model.loss_entropy()]);
                                 a_optimizer.zero_grad / backward / step                       (:610-621)

    as ONE CUDA graph (two forward/backward passes, two Adam launches).  `bilevel=False` is the warm-up schedule of
    search_lip_sync.py:325-326 (epochs < 15, core/function.py:57 `train`): the weight step only — its backward still
    produces the architecture gradients (alphas require grad), they are simply not applied.
    `entropy=True` adds the 2 * loss_entropy() term the reference switches on after epoch 70 (:612-616).
    Static inputs of the second batch: images2, par_lab2, edge_lab2, pose_gt2, pose_aux_gt2 (load2())."""

    def __init__(self, model, criterion_pose, criterion_par, optimizer, a_optimizer, batch, size=384, use_graph=True,
                 world_size=1, warmup=3, bilevel=True, entropy=False):
        super().__init__(model, criterion_pose, criterion_par, optimizer, batch, size, use_graph, world_size, warmup)
        self.a_opt = a_optimizer
        self.opts = [optimizer, a_optimizer]
        self.bilevel, self.entropy = bilevel, entropy
        self.a_flat = a_optimizer.use_flat_grads() if hasattr(a_optimizer, "use_flat_grads") else None
        self._in2 = [torch.zeros_like(t) for t in self._inputs()] if bilevel else None
        self.loss2 = torch.zeros((), device=self.images.device)

    def load2(self, images, par_lab, edge_lab, pose_gt, pose_aux_gt, non_blocking=True):
        for d, s in zip(self._in2, (images, par_lab, edge_lab, pose_gt, pose_aux_gt)):
            d.copy_(s, non_blocking=non_blocking)

    def input_bytes(self):
        n = super().input_bytes()
        return 2 * n if self.bilevel else n

    def _losses(self, images, par_lab, edge_lab, pose_gt, pose_aux_gt):
        pose, par = self.model(images)
        loss_par = self.cpar(par, [par_lab, edge_lab]).unsqueeze(0)
        loss_pose = self.cpose(pose, [pose_gt, pose_aux_gt]).unsqueeze(0)
        return loss_par + loss_pose

    def _pass(self, inputs, opt, alpha_pass):
        dev = self.images.device
        opt.zero_grad(set_to_none=True)
        if alpha_pass:
            self.opt.zero_grad(set_to_none=True)   # the alpha pass also back-propagates into the weights' slots: the
        else:                                      # reference leaves those (unused) gradients behind; so do we, zeroed
            self.a_opt.zero_grad(set_to_none=True)
        F_._arena.begin(dev)
        F_._state["defer_bn_counters"] = counters = []
        try:
            if self.packer is not None:
                self.packer.pack()
            losses = self._losses(*inputs)
            F_._state["defer_bn_counters"] = None
            if counters:
                torch._foreach_add_(counters, 1)
            if alpha_pass:
                if self.entropy:
                    losses = losses + 2 * self.model.loss_entropy()
                loss = 2 * losses.mean()
            else:
                loss = losses.mean()
            loss.backward()
        finally:
            F_._state["defer_bn_counters"] = None
            F_._arena.end()
            F_.WeightPacker.release()
        if self.world_size > 1:
            self._allreduce_grads(opt)
        opt.step()
        return loss.detach()

    def _step_body(self):
        self.loss.copy_(self._pass(self._inputs(), self.opt, False))
        if self.bilevel:
            self.loss2.copy_(self._pass(self._in2, self.a_opt, True))


class EvalStep:
    """Batched inference + on-GPU evaluation (BASELINE configs[4]; the pascal `validate_sync`, core/function_ppp.py:
    869-964): two eval-mode forwards (image and its mirror, :903-904), parsing logits resized to the label size and
    flip-averaged (:923-928, no left/right channel swap for pascal; swap_lr=True gives the LIP variant of
    core/function.py:927-939), int64 confusion histogram accumulated on the device (utils/utils.py:192-218), heat maps
    flip-averaged with the joint permutation in heat-map space (:957-958) and the PCK hit / valid counters of
    core/evaluate.py:43-99 — one CUDA graph per batch, nothing but the [C*C] + 2*[J] counters ever leaves the device."""
    FLIPPED_POSEIDX_PASCAL = [0, 1, 8, 9, 10, 11, 12, 13, 2, 3, 4, 5, 6, 7]       # function_ppp.py:905

    def __init__(self, model, batch, size=512, ignore_label=255, swap_lr=False, flipped_poseidx=None, use_graph=True):
        self.model = model.eval()
        dev = next(model.parameters()).device
        self.nc, self.nj = model._num_classes, model._num_joints
        self.size, self.ignore, self.swap_lr, self.use_graph = size, ignore_label, swap_lr, use_graph
        self.flip_idx = list(flipped_poseidx if flipped_poseidx is not None else
                             (self.FLIPPED_POSEIDX_PASCAL if self.nj == 14 else range(self.nj)))
        hs = size // 4
        self.images = torch.zeros(batch, 3, size, size, device=dev)
        self.par_lab = torch.zeros(batch, size, size, dtype=torch.int64, device=dev)
        self.pose_gt = torch.zeros(batch, self.nj, hs, hs, device=dev)
        self.hist = torch.zeros(self.nc * self.nc, dtype=torch.int64, device=dev)
        self.hit = torch.zeros(self.nj, dtype=torch.int64, device=dev)
        self.valid = torch.zeros(self.nj, dtype=torch.int64, device=dev)
        self.graph = None
        self.launches_per_step = None
        self.packer = F_.WeightPacker(model) if dev.type == "cuda" else None

    def load(self, images, par_lab, pose_gt, non_blocking=True):
        self.images.copy_(images, non_blocking=non_blocking)
        self.par_lab.copy_(par_lab, non_blocking=non_blocking)
        self.pose_gt.copy_(pose_gt, non_blocking=non_blocking)

    def input_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.images, self.par_lab, self.pose_gt))

    def reset(self):
        self.hist.zero_(), self.hit.zero_(), self.valid.zero_()

    @torch.no_grad()
    def _body(self):
        from .core import evaluate as ev
        from .utils import utils as U
        try:
            if self.packer is not None:
                self.packer.pack()
            pose, par = self.model(self.images)
            fpose, fpar = self.model(self.images.flip(3))
        finally:
            F_.WeightPacker.release()
        # the last-stage predictions of this batch stay readable after run() (static graph memory): tests evaluate the
        # oracle on exactly these tensors
        merged = U.tta_merge(par[-1][0], fpar[-1][0], (self.size, self.size), swap_lr=self.swap_lr)
        U.confusion_hist(self.par_lab, merged, self.nc, self.ignore, hist=self.hist)
        hm = ev.flip_average(pose[-1][0], fpose[-1][0], self.flip_idx)
        ev.pck_counts(hm, self.pose_gt, hit=self.hit, valid=self.valid)
        self.outputs = {"pose": pose[-1][0], "flip_pose": fpose[-1][0], "par": par[-1][0], "flip_par": fpar[-1][0],
                        "merged_par": merged, "merged_pose": hm}

    def prepare(self):
        if not self.use_graph or self.graph is not None:
            return
        hist, hit, valid = self.hist.clone(), self.hit.clone(), self.valid.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        c0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = _lib.launch_count() - c0
        self.hist.copy_(hist), self.hit.copy_(hit), self.valid.copy_(valid)   # the warm-up batch is not counted

    def run(self):
        if self.use_graph:
            if self.graph is None:
                self.prepare()
            self.graph.replay()
        else:
            c0 = _lib.launch_count()
            self._body()
            self.launches_per_step = _lib.launch_count() - c0

    def results(self):
        """(confusion matrix float64 [C, C] indexed [gt, pred], mean IoU, per-joint PCK in %, hit, valid) — the
        arithmetic of core/function.py:1026-1030 on the accumulated counters.  Synchronises."""
        import numpy as np
        cm = self.hist.cpu().numpy().astype(np.float64).reshape(self.nc, self.nc)
        pos, res, tp = cm.sum(1), cm.sum(0), np.diag(cm)
        iou = tp / np.maximum(1.0, pos + res - tp)
        hit, valid = self.hit.cpu().numpy(), self.valid.cpu().numpy()
        pck = 100.0 * hit / np.maximum(valid, 1)
        return cm, float(iou.mean()), pck, hit, valid


@torch.no_grad()
def inference(model, images):
    """Eval-mode forward (config 1 / config 5): returns (pose_list, par_list) like Network.forward."""
    model.eval()
    return model(images)
