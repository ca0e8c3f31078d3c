"""FusedAdam: torch.optim.Adam's arithmetic in ONE kernel launch per step (csrc/optim.cu).

Same constructor / param_groups interface as torch.optim.Adam (lr, betas, eps, weight_decay; no amsgrad), so it
drops into augment_lip_sync.py:210-212 (`Adam(param_dicts, LR)` + `add_param_group`).  The kernel walks a device
table of (param, grad, exp_avg, exp_avg_sq) pointers; the table is rebuilt only when a gradient's address
changes — never under CUDA-graph replay, where all addresses are static.
"""
import struct

import torch

from ._lib import call, fptr, i32, f32, stream

CHUNK = 16384


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._key = None
        self._tables = None       # (table, chunk_tensor, chunk_index) device tensors
        self._host = None         # pinned host copies awaiting upload (graph capture)
        self._step = None

    def _init_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _build(self, items, dev):
        rows, chunk_tensor, chunk_index = [], [], []
        for ti, (p, g, lr, wd) in enumerate(items):
            st = self._init_state(p)
            n = p.numel()
            rows.append(struct.pack("<QQQQqff", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                    st["exp_avg_sq"].data_ptr(), n, lr, wd))
            nch = (n + CHUNK - 1) // CHUNK
            chunk_tensor += [ti] * nch
            chunk_index += list(range(nch))
        # pageable host staging: no pinned allocation (a CUDA call) may happen while a graph is being captured
        tbl = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).clone()
        ct = torch.tensor(chunk_tensor, dtype=torch.int32)
        ci = torch.tensor(chunk_index, dtype=torch.int32)
        self._host = (tbl, ct, ci)
        self._tables = tuple(torch.empty_like(t, device=dev) for t in self._host)
        if not torch.cuda.is_current_stream_capturing():
            self.upload_tables()

    def upload_tables(self):
        """Copies the pointer tables to the device.  Called automatically in eager mode; after capturing a CUDA
        graph (or after changing a learning rate) call it once before the next replay."""
        if self._host is not None:
            for d, h in zip(self._tables, self._host):
                d.copy_(h, non_blocking=False)
            self._host = None

    def refresh_hyperparameters(self):
        """Rebuilds the table on the next step (e.g. after an lr scheduler changed group['lr'])."""
        self._key = None

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
        if self._step is None:
            self._step = torch.zeros((), dtype=torch.int64, device=dev)
            for p, _, _, _ in items:
                self.state[p]["step"] = self._step
        key = tuple((p.data_ptr(), g.data_ptr(), lr, wd) for p, g, lr, wd in items)
        if key != self._key:
            self._build(items, dev)
            self._key = key
        tbl, ct, ci = self._tables
        call("npp_adam_step", fptr(tbl), fptr(ct), fptr(ci), i32(ct.numel()), i32(CHUNK), fptr(self._step), f32(b1),
             f32(b2), f32(eps), stream())
        return None
