"""FusedAdam: torch.optim.Adam's arithmetic in ONE kernel launch per step (csrc/optim.cu).

Same constructor / param_groups interface as torch.optim.Adam (lr, betas, eps, weight_decay; no amsgrad), so it
drops into augment_lip_sync.py:210-212 (`Adam(param_dicts, LR)` + `add_param_group`).  The kernel walks a device
table of (param, grad, exp_avg, exp_avg_sq) pointers; the table is rebuilt only when a gradient's address
changes.  `use_flat_grads()` makes every gradient a persistent view into one flat fp32 buffer: addresses never
change (required for CUDA-graph capture: a table allocated while capturing would live in the graph's private
pool, whose blocks are reused by earlier graph nodes on every replay), zero_grad is one memset and the
data-parallel gradient all-reduce runs on the flat buffer without flatten / unflatten copies.
"""
import struct

import torch

from ._lib import call, fptr, i32, f32, stream

CHUNK = 16384


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._key = None          # addresses the device table was built for (params, grads, moments, step counters)
        self._hyper = None        # (lr, weight_decay) per table row as written into the device table
        self._tables = None       # (table, chunk_tensor, chunk_index) device tensors
        self._steps = None        # one int64 device counter per parameter (torch.optim.Adam's state['step'])
        self.flat_grads = None    # the flat gradient buffer once use_flat_grads() was called
        self._flat_views = None

    # ---- resume (augment_lip_sync.py:235 optimizer.load_state_dict) -----------------------------------------------
    def load_state_dict(self, state_dict):
        """torch.optim.Optimizer.load_state_dict, then: the loaded exp_avg / exp_avg_sq / step tensors are NEW
        tensors, so the device table (raw pointers) is rebuilt on the next step and the per-parameter step counters
        are re-seated in one int64 device vector holding the LOADED values (bias correction continues where the
        checkpoint left off).  Accepts state dicts written by torch.optim.Adam (float / CPU `step`)."""
        super().load_state_dict(state_dict)
        # torch moves / casts the loaded tensors only when needed, so they can still alias the caller's state dict (or
        # another optimizer's live state): the kernel updates them in place, hence private copies
        for st in self.state.values():
            for k, v in list(st.items()):
                if torch.is_tensor(v):
                    st[k] = v.clone()
        self._key = self._hyper = self._tables = None
        self._steps = None

    def __setstate__(self, state):
        super().__setstate__(state)
        self._key = self._hyper = self._tables = None
        self._steps = None

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if getattr(self, "_steps", None) is not None:
            self._key = self._hyper = None
            self._steps = None        # re-seated (values kept) on the next step

    def _seat_steps(self, dev):
        """One int64 device vector with a counter per parameter; state[p]['step'] are 0-dim views into it.  Counters
        that already exist (a loaded checkpoint, a previous seating) keep their values."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("FusedAdam: run one eager step (or load the checkpoint) before capturing a CUDA graph")
        allp = [p for group in self.param_groups for p in group["params"]]
        vals = []
        for p in allp:
            old = self.state[p].get("step") if p in self.state else None
            vals.append(int(float(old)) if old is not None else 0)
        self._steps = torch.tensor(vals, dtype=torch.int64).to(dev)
        for i, p in enumerate(allp):
            self.state[p]["step"] = self._steps[i]

    def _init_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def use_flat_grads(self):
        """Allocates ONE fp32 buffer holding the gradient of every parameter (16-byte aligned slots) and installs
        the slots as `p.grad`.  autograd then accumulates into them in place; zero_grad() zero-fills the buffer
        instead of dropping the gradients.  Parameters that take no part in the forward (SE_Block.bn at stride 1,
        operations.py:117) keep an all-zero gradient, for which the Adam update is exactly zero."""
        params = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        if not params:
            return None
        if self.flat_grads is not None and [id(p) for p, _ in self._flat_views] == [id(p) for p in params]:
            return self.flat_grads
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            if p.dtype != torch.float32 or p.device != dev or not p.is_contiguous():
                raise RuntimeError("FusedAdam.use_flat_grads needs contiguous fp32 parameters on one device")
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat_grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self._flat_views = [(p, self.flat_grads[o:o + p.numel()].view_as(p)) for p, o in zip(params, offs)]
        for p, v in self._flat_views:
            p.grad = v
            p._npp_grad_slot = v   # functional.grad_slot(): backward kernels accumulate here directly
        self._key = None
        return self.flat_grads

    def zero_grad(self, set_to_none=True):
        if self.flat_grads is None:
            return super().zero_grad(set_to_none=set_to_none)
        self.flat_grads.zero_()
        for p, v in self._flat_views:
            if p.grad is not v:
                p.grad = v

    def _build(self, items, dev):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError(
                "FusedAdam: gradient addresses changed while capturing a CUDA graph — call use_flat_grads() and "
                "run one eager step before the capture")
        rows, chunk_tensor, chunk_index = [], [], []
        for ti, (p, g, lr, wd) in enumerate(items):
            st = self._init_state(p)
            n = p.numel()
            rows.append(struct.pack("<QQQQqffQ", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                    st["exp_avg_sq"].data_ptr(), n, lr, wd, st["step"].data_ptr()))
            nch = (n + CHUNK - 1) // CHUNK
            chunk_tensor += [ti] * nch
            chunk_index += list(range(nch))
        # pageable host staging: no pinned allocation (a CUDA call) may happen while a graph is being captured
        tbl = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).clone()
        ct = torch.tensor(chunk_tensor, dtype=torch.int32)
        ci = torch.tensor(chunk_index, dtype=torch.int32)
        self._tables = tuple(t.to(dev) for t in (tbl, ct, ci))
        self._rows = rows

    def _current_hyper(self):
        """(lr, weight_decay) per table row, in table order (parameters with a gradient)."""
        return tuple((float(g["lr"]), float(g["weight_decay"])) for g in self.param_groups for p in g["params"]
                     if p.grad is not None)

    def sync_hyperparameters(self):
        """Writes the param groups' CURRENT lr / weight_decay into the device table IN PLACE (same device memory, so a
        captured CUDA graph picks the new values up on its next replay).  engine.TrainStep.run() calls this before
        every replay; it is a tuple comparison unless an lr scheduler (MultiStepLR, augment_lip_sync.py:213,249) or the
        user changed a group.  Returns True when the table was rewritten."""
        if self._tables is None or self._key is None:
            return False
        hyper = self._current_hyper()
        if hyper == self._hyper or len(hyper) != len(self._rows):
            return False
        rows = []
        for row, (lr, wd) in zip(self._rows, hyper):
            f = list(struct.unpack("<QQQQqffQ", row))
            f[5], f[6] = lr, wd
            rows.append(struct.pack("<QQQQqffQ", *f))
        host = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8)
        self._tables[0].copy_(host)      # stream-ordered before the next launch / graph replay on this stream
        self._rows, self._hyper = rows, hyper
        return True

    def refresh_hyperparameters(self):
        """Kept for callers of the first version: the table follows param_groups by itself now (step() and
        sync_hyperparameters() compare lr / weight_decay on every call)."""
        self.sync_hyperparameters()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("FusedAdam does not take a closure")
        items, dev = [], None
        b1 = b2 = eps = None
        for group in self.param_groups:
            gb1, gb2 = group["betas"]
            if b1 is None:
                b1, b2, eps = gb1, gb2, group["eps"]
            elif (gb1, gb2, group["eps"]) != (b1, b2, eps):
                raise RuntimeError("FusedAdam needs the same betas/eps in every param group")
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if g.dtype != torch.float32 or p.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 parameters and gradients")
                items.append((p, g, float(group["lr"]), float(group["weight_decay"])))
                dev = p.device
        if not items:
            return None
        if self._steps is None:
            self._seat_steps(dev)
        key = tuple((p.data_ptr(), g.data_ptr(), self._init_state(p)["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr(), self.state[p]["step"].data_ptr()) for p, g, _, _ in items)
        if key != self._key:
            self._build(items, dev)
            self._key = key
            self._hyper = tuple((lr, wd) for _, _, lr, wd in items)
        elif self._hyper != tuple((lr, wd) for _, _, lr, wd in items):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: lr / weight_decay changed while capturing; change them before the capture "
                                   "or between replays (TrainStep.run() syncs them)")
            self.sync_hyperparameters()
        tbl, ct, ci = self._tables
        call("npp_adam_step", fptr(tbl), i32(len(items)), fptr(ct), fptr(ci), i32(ct.numel()), i32(CHUNK), f32(b1),
             f32(b2), f32(eps), stream())
        return None
