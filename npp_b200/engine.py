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

    def _allreduce_grads(self):
        import torch.distributed as dist
        if self.flat_grads is not None:
            dist.all_reduce(self.flat_grads)
            self.flat_grads.div_(self.world_size)
            return
        params = [p for g in self.opt.param_groups for p in g["params"] if p.grad is not None]
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

    # ---- public ------------------------------------------------------------------------------------------
    def prepare(self):
        """Eager warm-up steps (cuTensorMap driver entry point, kernel attributes, optimizer state) and graph capture."""
        if self.graph is not None or not self.use_graph:
            if not self.use_graph and self.launches_per_step is None:
                c0 = _lib.launch_count()
                self._step_body()
                self.launches_per_step = _lib.launch_count() - c0
            return
        gc.collect()  # autograd graphs of earlier steps (AccumulateGrad nodes bound to another stream) must be gone
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

    def run(self):
        """Runs one step on whatever is in the static input buffers (or on the batch staged by prefetch()); returns
        the (device) loss scalar."""
        if self._staged:
            self._consume_staged()
        if self.use_graph:
            if self.graph is None:
                self.prepare()
            self.graph.replay()
        else:
            self._step_body()
        return self.loss


@torch.no_grad()
def inference(model, images):
    """Eval-mode forward (config 1 / config 5): returns (pose_list, par_list) like Network.forward."""
    model.eval()
    return model(images)
