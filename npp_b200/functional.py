"""Autograd bindings of the libnpp_b200 kernels.

Internal activation format ("internal tensor"): a torch tensor with logical shape [N, C, H, W],
channels_last strides (physical NHWC), dtype bf16 (product path) or fp32 (validation mode) and
C padded to a multiple of 8 with zero channels.  Every Function below launches hand-written
kernels on torch's current CUDA stream through the C ABI; torch only owns the memory.
"""
import ctypes
import os

import torch
from torch.autograd import Function

from . import _lib as L
from ._lib import call, view, stream, fptr, i32, i64, f32, f64, ref, NULL

_state = {"dtype": torch.bfloat16, "sync_bn": None, "defer_bn_counters": None,
          # kernel-selection switches (A/B timing, bisecting): NPP_NODE_STRIPED=1 -> striped-atomic totals instead of
          # partials + fold kernel for the BatchNorm backward sums of a node (measured on B200: 111.8 ms/step with 2
          # stripes, 110.1 with partials + fold — the apply kernel pays for folding the stripes — so OFF by default);
          # NPP_NODE_FUSED_FINALIZE=0 -> separate npp_bn_finalize launches
          "node_striped": os.environ.get("NPP_NODE_STRIPED", "0") != "0",
          "node_fused_finalize": os.environ.get("NPP_NODE_FUSED_FINALIZE", "1") != "0",
          # NPP_STEM_IM2COL=0 -> the 3-channel stems run through the generic 3x3 implicit-GEMM kernel
          "stem_im2col": os.environ.get("NPP_STEM_IM2COL", "1") != "0",
          # NPP_NODE_CAT_GRADS=0 -> the concat-slice gradient of a cell state is added by autograd (strided at::add)
          "node_cat_grads": os.environ.get("NPP_NODE_CAT_GRADS", "1") != "0",
          # NPP_BILINEAR_SEP=0 -> gather-form bilinear backward.  The separable two-pass form is the default: parity
          # green on a B200 (tests/test_gpu_zz_bilinear_sep.py), 32 registers per thread in both passes, and by the
          # traffic / instruction model of DESIGN.md section 4 about 1.7x cheaper than the gather form (237 us measured
          # at 512 channels 96^2 -> 24^2) — but it could not be TIMED before the round's GPU budget ran out
          "bilinear_sep": os.environ.get("NPP_BILINEAR_SEP", "1") != "0",
          # NPP_CONV_PAIR=1 -> 3x3 convolutions with 32 -> 32 channels run in pixel-pair form (round-2 candidate)
          "conv_pair": os.environ.get("NPP_CONV_PAIR", "0") != "0",
          # NPP_SE_BWD2=1 -> SE bottleneck backward in two kernels without weight-gradient atomics (round-2 candidate)
          "se_bwd2": os.environ.get("NPP_SE_BWD2", "0") != "0",
          # NPP_PACK_TILES=1 -> tile-transposing multi-tensor weight pack (round-2 candidate)
          "pack_tiles": os.environ.get("NPP_PACK_TILES", "0") != "0",
          # NPP_CE_BWD_SEP=1 -> cross-entropy backward as per-pixel gradient + separable bilinear backward instead of
          # the shared-memory-atomic kernel (round-2 candidate; core/criterion.py, csrc/loss.cu ce_grad_kernel)
          "ce_bwd_sep": os.environ.get("NPP_CE_BWD_SEP", "0") != "0",
          # NPP_TWO_STREAMS=0 -> single-stream schedule.  Default: the two task streams of the networks (cells1 / cells2, upsamples1 / upsamples2: the
          # same shapes, independent between the interaction points) are issued on two CUDA streams, so that inside a
          # captured step they are parallel graph branches: the small 12^2 / 24^2 / 48^2 launches of one stream fill
          # the SMs the other leaves idle (see TaskStreams below).  Measured on B200 (profiles/r02_bench_*_two_streams_*):
          # train 99.6 -> 90.4 ms, search 173.7 -> 159.1 ms, infer512 46.1 -> 41.3 ms per step
          "two_streams": os.environ.get("NPP_TWO_STREAMS", "1") != "0",
          # NPP_WGRAD_STREAM=1 -> the weight gradient of a dense convolution runs on a companion stream while the
          # data gradient runs on the layer's own stream (both only read dY); the own stream waits for the companion
          # before the backward node returns.  Measured with the two task streams already on: 89.18 vs 88.75 ms per
          # train step (profiles/r02_wgrad_stream_ab.txt) -> off by default
          "wgrad_stream": os.environ.get("NPP_WGRAD_STREAM", "0") == "1"}


def set_compute_dtype(dtype):
    """torch.bfloat16 (tcgen05 product path) or torch.float32 (fp32 validation mode)."""
    if dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("compute dtype must be torch.bfloat16 or torch.float32")
    _state["dtype"] = dtype


def get_compute_dtype():
    return _state["dtype"]


def pad8(c):
    return (int(c) + 7) // 8 * 8


def empty_internal(n, c, h, w, dtype, device):
    return torch.empty((n, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def is_internal(t):
    return (t.dim() == 4 and t.is_cuda and t.dtype in (torch.bfloat16, torch.float32) and t.shape[1] % 8 == 0
            and t.stride(1) == 1 and t.stride(3) % 8 == 0 and t.stride(2) % 8 == 0 and t.stride(0) % 8 == 0
            and t.data_ptr() % 16 == 0)


def as_internal_grad(g, like):
    """Gradients normally arrive in internal format from our own kernels; anything else (a grad
    produced by a torch op in user code) is re-laid-out once."""
    if g.dtype != like.dtype:
        g = g.to(like.dtype)
    if not is_internal(g) or g.stride(3) < g.shape[1]:
        g = g.contiguous(memory_format=torch.channels_last)
        if not is_internal(g):  # e.g. C == 1 edge cases cannot occur for padded tensors
            raise RuntimeError("cannot express gradient as an NHWC view: shape %s strides %s" % (g.shape, g.stride()))
    return g


def pad_vec(p, n, value=0.0):
    """Pads a 1-D parameter to n entries (autograd-aware); channel padding of internal tensors."""
    if p is None or p.numel() == n:
        return p
    return torch.cat([p, p.new_full((n - p.numel(),), value)])


# ------------------------------------------------------------------------------------------------
# per-step scratch: zero-initialised fp32 accumulators and direct parameter-gradient slots
# ------------------------------------------------------------------------------------------------
class ZeroArena:
    """One persistent fp32 buffer handing out the zero-initialised accumulators a step needs (BatchNorm statistic
    vectors, per-image SE sums, ...): ONE memset per step instead of one fill kernel per accumulator.  The step
    driver (engine.TrainStep) calls begin() before and end() after every step; slices are handed out in call order,
    anything that does not fit falls back to torch.zeros and grows the buffer for the next (eager) step."""

    def __init__(self):
        self.buf, self.cursor, self.need, self.active = None, 0, 0, False

    def begin(self, device):
        if self.buf is None or self.buf.numel() < self.need:
            if torch.device(device).type == "cuda" and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("ZeroArena must be sized by an eager step before CUDA-graph capture")
            self.buf = torch.zeros(max(self.need, 1 << 20), dtype=torch.float32, device=device)
        else:
            self.buf.zero_()
        self.cursor, self.need, self.active = 0, 0, True

    def end(self):
        self.active = False

    def reserve(self, device):
        """Grows the buffer to what the last step asked for.  engine.TrainStep.prepare() calls this between its eager
        warm-up step and the CUDA-graph capture (the buffer cannot be re-allocated while capturing)."""
        if self.buf is None or self.buf.numel() < self.need:
            self.buf = torch.zeros(max(self.need, 1 << 20), dtype=torch.float32, device=device)

    def take(self, n):
        n4 = (n + 3) // 4 * 4          # keep every slice 16-byte aligned (float4 loads)
        self.need += n4
        if self.buf is not None and self.cursor + n4 <= self.buf.numel():
            t = self.buf[self.cursor:self.cursor + n]
            self.cursor += n4
            return t
        return None


_arena = ZeroArena()


def zeros_f32(n, device):
    """fp32 zeros of n elements: an arena slice inside a TrainStep, torch.zeros otherwise."""
    if _arena.active:
        t = _arena.take(int(n))
        if t is not None:
            return t
    return torch.zeros(int(n), dtype=torch.float32, device=device)


# ------------------------------------------------------------------------------------------------
# two task streams on two CUDA streams
# ------------------------------------------------------------------------------------------------
class TaskStreams:
    """Issues the second task stream of a network (parsing: cells2, upsamples2, their stems) on a side CUDA stream.
    The two streams are independent between the interaction points of Network.forward, so inside a captured training
    step they become parallel graph branches; autograd runs every backward node on the stream of its forward, so the
    backward pass is parallel too.

        ts = TaskStreams(x)                 # main = the current stream
        ts.fork(x)                          # side waits for main; x (produced on main) may now be read on side
        with ts.side(): s3 = cell2(s2, s3)  # kernels + allocations of the block go to the side stream
        ts.join(s3)                         # main waits for side; s3 (produced on side) may now be read on main

    Memory safety: a tensor lives in the caching allocator's pool of the stream it was allocated on; every handle that
    crosses (fork / join arguments) is marked with record_stream for the other stream, so its block is not reused
    before that stream is done with it (during CUDA-graph capture: not before the capture ends).  Persistent scratch
    is per stream (_wgrad_workspace) or handed out once per step (ZeroArena); SyncBN exchanges of the side stream use
    their own communicator (distributed.enable_sync_bn creates two)."""
    _side = {}

    def __init__(self, ref, enabled=True):
        self.on = bool(enabled) and bool(_state.get("two_streams")) and ref.is_cuda
        if not self.on:
            return
        self.main = torch.cuda.current_stream()
        idx = ref.device.index
        if idx not in TaskStreams._side:
            TaskStreams._side[idx] = _register_stream(torch.cuda.Stream(device=ref.device), "side")
        self.b = TaskStreams._side[idx]

    @staticmethod
    def side_stream_of(device_index):
        return TaskStreams._side.get(device_index)

    def fork(self, *handles):
        if self.on:
            share(handles, self.b)
            self.b.wait_stream(self.main)

    def join(self, *handles):
        if self.on:
            share(handles, self.main)
            self.main.wait_stream(self.b)

    def side(self):
        import contextlib
        return torch.cuda.stream(self.b) if self.on else contextlib.nullcontext()


_WORKERS = {}     # (device index, "main" | "side") -> [worker streams]
_N_WORKERS = max(0, min(8, int(os.environ.get("NPP_BRANCH_STREAMS", "6"))))


def parallel_branches(fns):
    """Evaluates independent branches [fn() -> handle] on worker streams forked from the current stream and joins them
    before returning the results (the MixedOps that feed one node of a search cell, model_search_interact.py:352-356:
    3-6 chains of ~50 small kernels each, nothing between them to share).  Inside a captured step they are parallel
    graph branches; autograd keeps every branch's backward on its worker stream.  Inputs the branches read must have
    been produced on the current stream (they are marked for the workers via `inputs`); with SyncBN every worker
    stream has its own peer-memory communicator (distributed.enable_sync_bn) — branch j always runs on worker j % n, the
    same on every rank, so each communicator sees the same exchange sequence everywhere; over NCCL the branches stay on
    one stream.
    fns: list of (callable, inputs) pairs."""
    dev = torch.cuda.current_device() if torch.cuda.is_available() else None
    if (not _state.get("two_streams") or _N_WORKERS < 2 or len(fns) < 2 or dev is None
            or (_sync_group() is not None and not _state.get("peer_comms"))
            or not any(torch.is_tensor(t) and t.is_cuda for _, ins in fns for t in _flatten(ins))):
        return [fn() for fn, _ in fns]
    home = torch.cuda.current_stream()
    side = TaskStreams.side_stream_of(dev)
    key = (dev, "side" if (side is not None and home == side) else "main")   # one worker pool per task stream
    if key not in _WORKERS:
        _WORKERS[key] = [_register_stream(torch.cuda.Stream(), "worker:%s:%d" % (key[1], i)) for i in range(_N_WORKERS)]
    workers = _WORKERS[key]
    outs, used = [], set()
    for j, (fn, ins) in enumerate(fns):
        w = workers[j % len(workers)]
        share(ins, w)
        if j % len(workers) not in used:
            w.wait_stream(home)
            used.add(j % len(workers))
        with torch.cuda.stream(w):
            outs.append(fn())
    for j, o in enumerate(outs):
        share(o, home)
    for k in used:
        home.wait_stream(workers[k])
    return outs


def _flatten(h):
    if isinstance(h, (list, tuple)):
        for e in h:
            yield from _flatten(e)
    else:
        yield h


def share(h, stream):
    """record_stream on every tensor behind a handle (tensor, relu-only handle, Pending, list of handles)."""
    if h is None:
        return
    if isinstance(h, (list, tuple)):
        for e in h:
            share(e, stream)
        return
    if isinstance(h, Pending):
        share([h.y, h.stats, h.raw, h.relu] + list(h.spare), stream)
        return
    if torch.is_tensor(h) and h.is_cuda and h.numel():
        h.record_stream(stream)
        for attr in ("_npp_relu", "_npp_im2col"):
            t = getattr(h, attr, None)
            if isinstance(t, tuple):        # the im2col cache: (key, tensor)
                t = t[1]
            if torch.is_tensor(t) and t is not h and t.is_cuda:
                t.record_stream(stream)
                if getattr(t, "_npp_home", None) is not None:
                    t._npp_ok = set(getattr(t, "_npp_ok", ())) | {stream.cuda_stream}


def grad_slot(p):
    """The persistent gradient slot of a parameter when the optimizer installed one (FusedAdam.use_flat_grads):
    backward kernels accumulate into it directly and autograd gets no gradient to add (one kernel launch per
    parameter less).  None when gradients are ordinary autograd-managed tensors."""
    if p is None or not torch.is_grad_enabled():
        return None
    s = getattr(p, "_npp_grad_slot", None)
    return s if (s is not None and p.grad is s) else None


# ------------------------------------------------------------------------------------------------
# layout conversion at the module edge
# ------------------------------------------------------------------------------------------------
class _ToInternal(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        n, c, h, w = x.shape
        ctx.c = c
        xs = x.contiguous().float()
        y = empty_internal(n, pad8(c), h, w, dtype, x.device)
        call("npp_nchw_to_nhwc", fptr(xs), i32(c), ref(view(y)), i32(L.dtype_code(y)), stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_internal_grad(dy, dy)
        n, _, h, w = dy.shape
        dx = torch.empty((n, ctx.c, h, w), dtype=torch.float32, device=dy.device)
        call("npp_nhwc_to_nchw", ref(view(dy)), fptr(dx), i32(ctx.c), i32(L.dtype_code(dy)), stream())
        return dx, None


class _FromInternal(Function):
    @staticmethod
    def forward(ctx, x, c):
        n, cp, h, w = x.shape
        ctx.cp, ctx.dtype = cp, x.dtype
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        call("npp_nhwc_to_nchw", ref(view(x)), fptr(y), i32(c), i32(L.dtype_code(x)), stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w = dy.shape
        dys = dy.contiguous().float()
        dx = empty_internal(n, ctx.cp, h, w, ctx.dtype, dy.device)
        call("npp_nchw_to_nhwc", fptr(dys), i32(c), ref(view(dx)), i32(L.dtype_code(dx)), stream())
        return dx, None


def to_internal(x, dtype=None):
    """NCHW fp32 (any torch layout) -> internal; internal tensors pass through; a Pending BatchNorm output is
    normalised (one fused pass) and cached."""
    if isinstance(x, Pending):
        return finish(x)
    if is_internal(x) and (dtype is None or x.dtype == dtype):
        return x
    return _ToInternal.apply(x, dtype or _state["dtype"])


def from_internal(x, c=None):
    """internal -> contiguous NCHW fp32 with the first c channels (drops channel padding)."""
    x = check_raw(to_internal(x), "from_internal")
    return _FromInternal.apply(x, int(c if c is not None else x.shape[1]))


# ------------------------------------------------------------------------------------------------
# ReLU
# ------------------------------------------------------------------------------------------------
class _ReluFn(Function):
    @staticmethod
    def forward(ctx, x):
        y = empty_internal(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.dtype, x.device)
        call("npp_relu_fwd", ref(view(x)), ref(view(y)), i32(L.dtype_code(x)), stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = as_internal_grad(dy, y)
        dx = torch.empty_like(y)
        call("npp_relu_bwd", ref(view(y)), ref(view(dy)), ref(view(dx)), i32(L.dtype_code(y)), stream())
        return dx


def relu(x):
    """nn.ReLU; the result is cached on the input so that the several primitives of a cell that read the same
    state (each starts with its own nn.ReLU, operations.py:76,212) share one pass — and normally it already
    exists: the kernel that produced the state wrote relu(state) next to it (node())."""
    if isinstance(x, Pending):
        materialize(x)
        return x.spare.pop(0) if x.spare else x.relu
    y = relu_of(x)
    if y is None:
        y = _ReluFn.apply(x)
        y._npp_is_relu = True
        if _state.get("two_streams") and y.is_cuda:
            y._npp_home = torch.cuda.current_stream().cuda_stream    # a lazily created cache entry: see relu_of
        x._npp_relu = y
    return y


def relu_of(x):
    """relu(x) if it already exists (written by the producer of x, or x is itself a ReLU output), else None.
    With two task streams a relu that was created lazily by a consumer on ONE stream is not visible to the other
    (no ordering, no allocator mark) unless the handle has crossed a TaskStreams fork / join since (share())."""
    if getattr(x, "_npp_is_relu", False):
        return x
    y = getattr(x, "_npp_relu", None)
    if y is not None:
        home = getattr(y, "_npp_home", None)
        if home is not None:
            cur = torch.cuda.current_stream().cuda_stream
            if cur != home and cur not in getattr(y, "_npp_ok", ()):
                return None
    return y


def check_raw(x, what):
    """Raises when `x` is a handle that only carries relu(state) (the producer was told nobody reads the raw
    state) but `what` needs the raw values."""
    if getattr(x, "_npp_relu_only", False):
        raise RuntimeError("%s needs the raw state, but its producer only materialised relu(state)" % what)
    return x


# ------------------------------------------------------------------------------------------------
# dense convolution
# ------------------------------------------------------------------------------------------------
def _wgrad_workspace(device):
    """One persistent fp32 workspace per (device, stream) for the split-K partial tiles of npp_conv2d_wgrad_ws (every
    wgrad call of a stream reuses it; stream order keeps the calls apart)."""
    key = ("wgrad_ws", torch.device(device).index, _stream_role())
    ws = _state.get(key)
    if ws is None:
        L.lib().npp_conv2d_wgrad_workspace_bytes.restype = ctypes.c_int64
        nbytes = int(L.lib().npp_conv2d_wgrad_workspace_bytes())
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
        _state[key] = ws
    return ws


_COMPANIONS = {}
_ROLES = {}       # cuda_stream handle of every persistent helper stream (task side stream, workers, companions) -> role


def _stream_role():
    """'main' for the caller's stream (default, warm-up or graph-capture stream), else the name of the persistent
    helper stream the code is running on.  Per-stream scratch (wgrad workspace, companions) is keyed by role, so the
    warm-up and the capture pass of a step share it."""
    return _ROLES.get(torch.cuda.current_stream().cuda_stream, "main")


def _register_stream(s, role):
    _ROLES[s.cuda_stream] = role
    return s


def _companion_stream():
    """The wgrad companion of the current stream (one per stream that runs convolutions backward)."""
    key = (torch.cuda.current_device(), _stream_role())
    c = _COMPANIONS.get(key)
    if c is None:
        c = _COMPANIONS[key] = _register_stream(torch.cuda.Stream(), "companion:" + key[1])
    return c


def conv_out_size(size, k, stride, pad, dil, off=0):
    return (size - off + 2 * pad - dil * (k - 1) - 1) // stride + 1


def _conv_work(n, ho, wo, cout, cin, kh, kw, x, y):
    """(algorithmic FLOPs, algorithmic bytes) of one conv direction: 2*N*Ho*Wo*Cout*Cin*kh*kw; unique input +
    output elements + weights once."""
    flops = 2.0 * n * ho * wo * cout * cin * kh * kw
    esz = x.element_size()
    byts = float(x.shape[0] * x.shape[2] * x.shape[3] * cin * esz + n * ho * wo * cout * esz + cout * cin * kh * kw * esz)
    return flops, byts


class _ConvFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, dil, hoff, woff, want_stats, slots, packed):
        n, cx, h, w = x.shape
        cout, cin, kh, kw = weight.shape
        if cx != pad8(cin):
            raise RuntimeError("conv input has %d channels, weight expects %d (padded %d)" % (cx, cin, pad8(cin)))
        cop = pad8(cout)
        ho, wo = conv_out_size(h, kh, stride, pad, dil, hoff), conv_out_size(w, kw, stride, pad, dil, woff)
        code = L.dtype_code(x)
        bf16 = code == L.NPP_BF16
        need_wt = bf16 and ctx.needs_input_grad[0]
        if packed is not None and bf16 and packed[0].numel() == cop * kh * kw * cx:
            wp, wt = packed        # packed for the whole model by one launch at the start of the step (WeightPacker)
        else:
            w32 = weight.detach().contiguous()
            wp = torch.empty(cop * kh * kw * cx, dtype=x.dtype, device=x.device)
            wt = torch.empty_like(wp) if need_wt else None
            call("npp_pack_weight", fptr(w32), fptr(wp), fptr(wt), i32(cout), i32(kh * kw), i32(cin), i32(cop), i32(cx),
                 i32(code), stream())
        bp = None
        if bias is not None:
            bp = pad_vec(bias.detach().float(), cop).contiguous()
        y = empty_internal(n, cop, ho, wo, x.dtype, x.device)
        stats = None
        if bf16:
            if want_stats:
                stats = zeros_f32(2 * cop, x.device)
            call("npp_conv2d_fwd", ref(view(x)), fptr(wp), fptr(bp), ref(view(y)), i32(kh), i32(kw), i32(stride),
                 i32(pad), i32(dil), i32(hoff), i32(woff), fptr(stats), stream(),
                 work=_conv_work(n, ho, wo, cout, cin, kh, kw, x, y), keep=(x, wp, bp, y, stats))
        else:
            call("npp_conv2d_direct_fwd", ref(view(x)), fptr(wp), fptr(bp), ref(view(y)), i32(kh), i32(kw),
                 i32(stride), i32(pad), i32(dil), i32(hoff), i32(woff), i32(code), stream())
        ctx.cfg = (stride, pad, dil, hoff, woff, cout, cin, kh, kw, bias is not None)
        ctx.slots = slots if slots is not None else (None, None)
        ctx.work = _conv_work(n, ho, wo, cout, cin, kh, kw, x, y)
        ctx.save_for_backward(x, wt if bf16 else wp)
        if stats is None:
            stats = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        x, wmat = ctx.saved_tensors
        stride, pad, dil, hoff, woff, cout, cin, kh, kw, has_bias = ctx.cfg
        dy = as_internal_grad(dy, x)
        code = L.dtype_code(x)
        bf16 = code == L.NPP_BF16
        dx = dw = db = None
        wslot, bslot = ctx.slots
        # wgrad on the companion stream of this layer's stream, concurrently with dgrad (issued first so that it is
        # already running when the dgrad kernel arrives); joined before this function returns
        side = None
        if (bf16 and ctx.needs_input_grad[0] and ctx.needs_input_grad[1] and _state.get("wgrad_stream")
                and _state.get("two_streams") and L._trace is None and L._prof is None):
            side = _companion_stream()
        if ctx.needs_input_grad[1] and side is not None:
            dw = wslot if wslot is not None else torch.zeros((cout, cin, kh, kw), dtype=torch.float32, device=x.device)
            own = torch.cuda.current_stream()
            side.wait_stream(own)
            with torch.cuda.stream(side):
                ws = _wgrad_workspace(x.device)
                call("npp_conv2d_wgrad_ws", ref(view(x)), ref(view(dy)), fptr(dw), i32(cout), i32(cin), i32(kh), i32(kw),
                     i32(stride), i32(pad), i32(dil), i32(hoff), i32(woff), fptr(ws), i64(ws.numel() * 4), stream(),
                     work=ctx.work, keep=(x, dy, dw))
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if bf16:
                call("npp_conv2d_dgrad", ref(view(dy)), fptr(wmat), ref(view(dx)), i32(kh), i32(kw), i32(stride),
                     i32(pad), i32(dil), i32(hoff), i32(woff), stream(), work=ctx.work, keep=(dy, wmat, dx))
            else:
                call("npp_conv2d_direct_dgrad", ref(view(dy)), fptr(wmat), ref(view(dx)), i32(kh), i32(kw),
                     i32(stride), i32(pad), i32(dil), i32(hoff), i32(woff), i32(code), stream())
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        if ctx.needs_input_grad[1] and side is None:
            # wgrad accumulates (+=): into the parameter's own gradient slot when there is one, else into zeros
            dw = wslot if wslot is not None else torch.zeros((cout, cin, kh, kw), dtype=torch.float32, device=x.device)
            if bf16:
                ws = _wgrad_workspace(x.device)
                call("npp_conv2d_wgrad_ws", ref(view(x)), ref(view(dy)), fptr(dw), i32(cout), i32(cin), i32(kh), i32(kw),
                     i32(stride), i32(pad), i32(dil), i32(hoff), i32(woff), fptr(ws), i64(ws.numel() * 4), stream(),
                     work=ctx.work, keep=(x, dy, dw))
            else:
                call("npp_conv2d_direct_wgrad", ref(view(x)), ref(view(dy)), fptr(dw), i32(cout), i32(cin), i32(kh),
                     i32(kw), i32(stride), i32(pad), i32(dil), i32(hoff), i32(woff), i32(code), stream())
            if wslot is not None:
                dw = None
        if has_bias and ctx.needs_input_grad[2]:
            if bslot is not None and dy.shape[1] == cout:
                call("npp_colsum", ref(view(dy)), fptr(bslot), i32(code), stream())
            else:
                dbp = zeros_f32(dy.shape[1], x.device)
                call("npp_colsum", ref(view(dy)), fptr(dbp), i32(code), stream())
                db = dbp[:cout]
        return dx, dw, db, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# pixel-pair form of the C = 32 3x3 convolutions (round-2 candidate, NPP_CONV_PAIR=1; csrc/conv_tcgen05.cu pair_weight)
# ------------------------------------------------------------------------------------------------
def _pair_ok(t):
    """dense NHWC [N, 32, H, W] with an even width: can be re-read as [N, 64, H, W/2]."""
    if t.dim() != 4 or t.shape[1] != 32 or t.shape[3] % 2:
        return False
    n, c, h, w = t.shape
    return tuple(t.stride()) == (h * w * c, 1, w * c, c)


def _pair_alias(t):
    n, c, h, w = t.shape
    return torch.empty(0, dtype=t.dtype, device=t.device).set_(
        t.untyped_storage(), t.storage_offset(), (n, 2 * c, h, w // 2), (h * w * c, 1, w * c, 2 * c))


class _ConvPairFn(Function):
    """3x3 / stride 1 / pad 1 convolution with 32 -> 32 channels computed as a 64 -> 64 convolution on the same
    memory read as super-pixels of two neighbouring pixels: half as many (full 128-byte) TMA rows, K = N = 64 instead
    of a half-empty 64-wide K block and N = 32.  The weight matrix, the BatchNorm sums and the weight gradient are
    mapped between the two forms by npp_pack_weight_pair / a fold of the [2][2][32] sums / npp_fold_pair_wgrad."""

    @staticmethod
    def forward(ctx, x, weight, want_stats, wslot, packed):
        n, c, h, w = x.shape
        if packed is not None:
            wp, wt = packed
        else:
            wp = torch.empty(64 * 9 * 64, dtype=x.dtype, device=x.device)
            wt = torch.empty_like(wp)
            call("npp_pack_weight_pair", fptr(weight.detach().contiguous()), fptr(wp), fptr(wt), stream())
        y = empty_internal(n, c, h, w, x.dtype, x.device)
        xv, yv = _pair_alias(x), _pair_alias(y)
        stats64 = zeros_f32(2 * 64, x.device) if want_stats else None
        call("npp_conv2d_fwd", ref(view(xv)), fptr(wp), NULL, ref(view(yv)), i32(3), i32(3), i32(1), i32(1), i32(1), i32(0),
             i32(0), fptr(stats64), stream(), work=_conv_work(n, h, w, 32, 32, 3, 3, x, y), keep=(x, wp, y, stats64))
        ctx.wslot = wslot
        ctx.work = _conv_work(n, h, w, 32, 32, 3, 3, x, y)
        ctx.save_for_backward(x, wt)
        if stats64 is not None:
            stats = stats64.view(2, 2, 32).sum(1).reshape(64)   # (moment, pixel parity, channel) -> (moment, channel)
        else:
            stats = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        x, wt = ctx.saved_tensors
        dy = as_internal_grad(dy, x)
        if not _pair_ok(dy):
            dy = dy.contiguous(memory_format=torch.channels_last)
        xv, dyv = _pair_alias(x), _pair_alias(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("npp_conv2d_dgrad", ref(view(dyv)), fptr(wt), ref(view(_pair_alias(dx))), i32(3), i32(3), i32(1), i32(1),
                 i32(1), i32(0), i32(0), stream(), work=ctx.work, keep=(dy, wt, dx))
        if ctx.needs_input_grad[1]:
            dwp = zeros_f32(64 * 64 * 9, x.device)
            ws = _wgrad_workspace(x.device)
            call("npp_conv2d_wgrad_ws", ref(view(xv)), ref(view(dyv)), fptr(dwp), i32(64), i32(64), i32(3), i32(3), i32(1),
                 i32(1), i32(1), i32(0), i32(0), fptr(ws), i64(ws.numel() * 4), stream(), work=ctx.work, keep=(x, dy, dwp))
            dw = ctx.wslot if ctx.wslot is not None else torch.zeros((32, 32, 3, 3), dtype=torch.float32, device=x.device)
            call("npp_fold_pair_wgrad", fptr(dwp), fptr(dw), stream(), keep=(dwp, dw))
            if ctx.wslot is not None:
                dw = None
        return dx, dw, None, None, None


def im2col3x3_c3(x, stride, pad):
    """[N, 3(+5 pad), H, W] internal bf16 image -> [N, 32, Ho, Wo] with channel ci*9 + r*3 + s = tap (r, s) of input
    channel ci (npp_im2col3x3_c3).  The result is cached on the image tensor: both task streams' stems
    (model_augment.py:244-272) read the same input."""
    key = (int(stride), int(pad))
    hit = getattr(x, "_npp_im2col", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    n, _, h, w = x.shape
    ho, wo = conv_out_size(h, 3, stride, pad, 1), conv_out_size(w, 3, stride, pad, 1)
    y = empty_internal(n, 32, ho, wo, x.dtype, x.device)
    call("npp_im2col3x3_c3", ref(view(x)), ref(view(y)), i32(stride), i32(pad), stream())
    x._npp_im2col = (key, y)
    return y


def conv2d(x, weight, bias=None, stride=1, pad=0, dil=1, hoff=0, woff=0, want_stats=False, slot_of=None):
    """Dense conv on an internal tensor.  Returns (y, stats) where stats is the fused per-channel
    (sum, sum of squares) of y from the tcgen05 epilogue (empty when not requested / fp32 mode).
    slot_of: the parameter `weight` is a reshaped view of (its gradient slot, viewed the same way, receives dW)."""
    wslot = grad_slot(weight) if slot_of is None else grad_slot(slot_of)
    if wslot is not None and slot_of is not None:
        wslot = wslot.view(weight.shape)
    if (_state.get("conv_pair", False) and bias is None and tuple(weight.shape) == (32, 32, 3, 3) and stride == 1
            and pad == 1 and dil == 1 and not hoff and not woff and x.dtype == torch.bfloat16 and _pair_ok(x)):
        packed = getattr(weight, "_npp_packed_pair", None) if _state.get("packed_weights") else None
        return _ConvPairFn.apply(x, weight, bool(want_stats), wslot, packed)
    slots = (wslot, grad_slot(bias))
    packed = getattr(weight, "_npp_packed", None) if _state.get("packed_weights") else None
    return _ConvFn.apply(x, weight, bias, int(stride), int(pad), int(dil), int(hoff), int(woff), bool(want_stats),
                         slots if (slots[0] is not None or slots[1] is not None) else None, packed)


class WeightPacker:
    """Packs every dense-conv weight of a model (fp32 OIHW masters -> bf16 OHWI + transposed copies) with ONE launch
    per step instead of one per convolution.  The packed buffers are persistent and hang off the parameters
    (`_npp_packed`); conv2d() uses them while `_state["packed_weights"]` is set (engine.TrainStep brackets its step
    body with pack() / release())."""
    CHUNK = 16384

    def __init__(self, model):
        import struct
        convs = [m for m in model.modules()
                 if isinstance(m, torch.nn.Conv2d) and m.groups == 1 and m.weight.is_cuda and m.weight.dtype == torch.float32]
        self.params = [m.weight for m in convs]
        self.bufs = []
        rows, ct, ci = [], [], []
        for ti, w in enumerate(self.params):
            cout, cin, kh, kw = w.shape
            cop, cip = pad8(cout), pad8(cin)
            n = cop * kh * kw * cip
            prev = getattr(w, "_npp_packed", None)
            if prev is not None and prev[0].numel() == n and prev[0].device == w.device:
                wp, wt = prev      # another packer of the same model (e.g. an eager and a graphed TrainStep) owns them too
            else:
                wp = torch.empty(n, dtype=torch.bfloat16, device=w.device)
                wt = torch.empty(n, dtype=torch.bfloat16, device=w.device)
                w._npp_packed = (wp, wt)
            self.bufs.append((wp, wt))   # the device table holds raw pointers: keep the buffers alive
            rows.append(struct.pack("<QQQiiiiii", w.data_ptr(), wp.data_ptr(), wt.data_ptr(), cout, kh * kw, cin, cop, cip, 0))
            nch = (n + self.CHUNK - 1) // self.CHUNK
            ct += [len(rows) - 1] * nch
            ci += list(range(nch))
            if _state.get("conv_pair", False) and (cout, cin, kh, kw) == (32, 32, 3, 3):
                # pixel-pair layout of the same weight (table flag pad == 1), see _ConvPairFn
                prev = getattr(w, "_npp_packed_pair", None)
                if prev is not None and prev[0].device == w.device:
                    pp, pt = prev
                else:
                    pp = torch.empty(64 * 9 * 64, dtype=torch.bfloat16, device=w.device)
                    pt = torch.empty_like(pp)
                    w._npp_packed_pair = (pp, pt)
                self.bufs.append((pp, pt))
                rows.append(struct.pack("<QQQiiiiii", w.data_ptr(), pp.data_ptr(), pt.data_ptr(), 32, 9, 32, 64, 64, 1))
                nch = (64 * 9 * 64 + self.CHUNK - 1) // self.CHUNK
                ct += [len(rows) - 1] * nch
                ci += list(range(nch))
        dev = self.params[0].device if self.params else None
        self.n = len(rows)
        self.tiles = None
        if self.n:
            self.table = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).clone().to(dev)
            self.chunk_tensor = torch.tensor(ct, dtype=torch.int32).to(dev)
            self.chunk_index = torch.tensor(ci, dtype=torch.int32).to(dev)
            self.ptrs = [w.data_ptr() for w in self.params]
            if _state.get("pack_tiles", False):
                # tile form (round-2 candidate): one block per 32 x 32 x taps tile of every plain (non-pair) row; rows
                # the tile kernel cannot take (more than 9 taps, pair layout) stay on the element-wise kernel
                tt, tix, ect, eci = [], [], [], []
                for ri, row in enumerate(rows):
                    _, _, _, cout, taps, cin, cop, cip, flag = struct.unpack("<QQQiiiiii", row)
                    if flag == 0 and taps <= 9:
                        nt = ((cop + 31) // 32) * ((cip + 31) // 32)
                        tt += [ri] * nt
                        tix += list(range(nt))
                    else:
                        tot = 64 * 9 * 64 if flag == 1 else cop * taps * cip
                        nch = (tot + self.CHUNK - 1) // self.CHUNK
                        ect += [ri] * nch
                        eci += list(range(nch))
                self.tiles = (torch.tensor(tt, dtype=torch.int32).to(dev), torch.tensor(tix, dtype=torch.int32).to(dev))
                self.chunk_tensor = torch.tensor(ect, dtype=torch.int32).to(dev)
                self.chunk_index = torch.tensor(eci, dtype=torch.int32).to(dev)

    def pack(self):
        if not self.n or _state["dtype"] != torch.bfloat16:
            return
        if [w.data_ptr() for w in self.params] != self.ptrs:
            raise RuntimeError("WeightPacker: a parameter was re-allocated; build a new packer")
        if self.tiles is not None and self.tiles[0].numel():
            call("npp_pack_weights_tiles", fptr(self.table), i32(self.n), fptr(self.tiles[0]), fptr(self.tiles[1]),
                 i32(self.tiles[0].numel()), stream())
        if self.chunk_tensor.numel():
            call("npp_pack_weights_multi", fptr(self.table), i32(self.n), fptr(self.chunk_tensor), fptr(self.chunk_index),
                 i32(self.chunk_tensor.numel()), i32(self.CHUNK), stream())
        _state["packed_weights"] = True

    @staticmethod
    def release():
        _state["packed_weights"] = False


# ------------------------------------------------------------------------------------------------
# depthwise convolution (+ fused leading ReLU)
# ------------------------------------------------------------------------------------------------
class _DwConvFn(Function):
    @staticmethod
    def forward(ctx, x, weight, stride, pad, dil, relu_in, wslot):
        n, c, h, w = x.shape
        k = weight.shape[-1]
        ctx.wslot = wslot
        if weight.shape[0] != c or weight.shape[1] != 1:
            raise RuntimeError("depthwise weight %s does not match %d channels" % (tuple(weight.shape), c))
        ho, wo = conv_out_size(h, k, stride, pad, dil), conv_out_size(w, k, stride, pad, dil)
        y = empty_internal(n, c, ho, wo, x.dtype, x.device)
        w32 = weight.detach().contiguous()
        call("npp_dwconv_fwd", ref(view(x)), fptr(w32), ref(view(y)), i32(k), i32(stride), i32(pad), i32(dil),
             i32(relu_in), i32(L.dtype_code(x)), stream())
        ctx.cfg = (k, stride, pad, dil, relu_in)
        ctx.save_for_backward(x, w32)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w32 = ctx.saved_tensors
        k, stride, pad, dil, relu_in = ctx.cfg
        dy = as_internal_grad(dy, x)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = ctx.wslot if ctx.wslot is not None else torch.zeros_like(w32)
        call("npp_dwconv_bwd", ref(view(x)), fptr(w32), ref(view(dy)), ref(view(dx)) if dx is not None else NULL,
             fptr(dw), i32(k), i32(stride), i32(pad), i32(dil), i32(relu_in), i32(L.dtype_code(x)), stream())
        return dx, (None if ctx.wslot is not None else dw), None, None, None, None, None


def dwconv2d(x, weight, stride, pad, dil, relu_in):
    return _DwConvFn.apply(x, weight, int(stride), int(pad), int(dil), int(bool(relu_in)), grad_slot(weight))


# ------------------------------------------------------------------------------------------------
# BatchNorm (training: batch statistics, optionally all-reduced across ranks = SyncBN)
# ------------------------------------------------------------------------------------------------
def _sync_group():
    g = _state["sync_bn"]
    if g is None:
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(g if g is not True else None) == 1:
        return None
    return g


def _allreduce_sum(t, t2=None):
    """In-place SUM over the SyncBN group of one or two fp32 statistic vectors; returns the group size.  On CUDA with
    a peer communicator (distributed.enable_sync_bn) this is ONE single-block kernel reading the peers' staging
    buffers over NVLink (csrc/peer.cu) instead of an NCCL collective per vector."""
    import torch.distributed as dist
    g = _state["sync_bn"]
    grp = None if g is True else g
    comm = _state.get("peer_comm")
    if comm is not None and t.is_cuda:
        # one communicator (= one device-side sequence counter) per stream that issues exchanges
        role = _stream_role()
        if role == "side" and _state.get("peer_comm_side") is not None:
            comm = _state["peer_comm_side"]
        elif role.startswith("worker:"):
            comm = (_state.get("peer_comms") or {}).get(role)
            if comm is None:
                raise RuntimeError("SyncBN exchange on worker stream %r without a communicator" % role)
        comm.allreduce(t, t2)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=grp)
        if t2 is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.SUM, group=grp)
    return dist.get_world_size(grp)


def _sync_world():
    import torch.distributed as dist
    g = _state["sync_bn"]
    return dist.get_world_size(None if g is True else g)


class _BNTrainFn(Function):
    @staticmethod
    def forward(ctx, x, stats, gamma, beta, running_mean, running_var, momentum, eps, relu, residual):
        n, c, h, w = x.shape
        code = L.dtype_code(x)
        dev = x.device
        if stats is None or stats.numel() == 0:
            stats = torch.zeros(2 * c, dtype=torch.float32, device=dev)
            call("npp_bn_stats", ref(view(x)), fptr(stats), i32(code), stream())
        count = float(n * h * w)
        sync = _sync_group() is not None
        if sync:
            count *= _allreduce_sum(stats)
        coef = torch.empty(4 * c, dtype=torch.float32, device=dev)  # scale | shift | mean | invstd
        scale, shift, mean, invstd = coef[:c], coef[c:2 * c], coef[2 * c:3 * c], coef[3 * c:]
        c_run = running_mean.numel() if running_mean is not None else 0
        call("npp_bn_finalize", fptr(stats), f64(count), fptr(gamma), fptr(beta), fptr(running_mean),
             fptr(running_var), f32(momentum), f32(eps), fptr(scale), fptr(shift), fptr(mean), fptr(invstd), i32(c),
             i32(c_run), stream())
        y = empty_internal(n, c, h, w, x.dtype, dev)
        call("npp_bn_apply", ref(view(x)), fptr(scale), fptr(shift), ref(view(residual)) if residual is not None else NULL,
             i32(relu), ref(view(y)), i32(code), stream())
        ctx.relu, ctx.count, ctx.sync, ctx.has_res = relu, count, sync, residual is not None
        ctx.save_for_backward(x, gamma, mean, invstd, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, invstd, y = ctx.saved_tensors
        dy = as_internal_grad(dy, x)
        c = x.shape[1]
        code = L.dtype_code(x)
        sums = torch.zeros(2 * c, dtype=torch.float32, device=x.device)
        ymask = ref(view(y)) if ctx.relu else NULL
        call("npp_bn_bwd_reduce", ref(view(dy)), ref(view(x)), ymask, fptr(mean), fptr(invstd), fptr(sums), i32(code),
             stream())
        dbeta, dgamma = sums[:c], sums[c:]
        if ctx.sync:
            dbeta, dgamma = dbeta.clone(), dgamma.clone()  # parameter grads stay local (DDP averages them)
            _allreduce_sum(sums)
        dx = torch.empty_like(x)
        call("npp_bn_bwd_apply", ref(view(dy)), ref(view(x)), ymask, fptr(gamma), fptr(mean), fptr(invstd), fptr(sums),
             f64(ctx.count), ref(view(dx)), i32(code), stream())
        dres = None
        if ctx.has_res:
            if ctx.relu:
                raise RuntimeError("residual + relu fusion has no backward")
            dres = dy
        return (dx, None, dgamma if ctx.needs_input_grad[2] else None, dbeta if ctx.needs_input_grad[3] else None,
                None, None, None, None, None, dres)


class _AffineFn(Function):
    """y = x*scale[c] + shift[c] (+relu): eval-mode BatchNorm with running statistics."""

    @staticmethod
    def forward(ctx, x, scale, shift, relu):
        y = empty_internal(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.dtype, x.device)
        call("npp_bn_apply", ref(view(x)), fptr(scale), fptr(shift), NULL, i32(relu), ref(view(y)),
             i32(L.dtype_code(x)), stream())
        if x.requires_grad:
            raise RuntimeError("eval-mode BatchNorm backward is not implemented; run under torch.no_grad()")
        return y

    @staticmethod
    def backward(ctx, dy):
        raise RuntimeError("eval-mode BatchNorm backward is not implemented")


def batch_norm(x, stats, gamma, beta, running_mean, running_var, training, momentum, eps, relu=False, residual=None):
    c = x.shape[1]
    gp = pad_vec(gamma, c, 1.0)
    bp = pad_vec(beta, c, 0.0)
    if training:
        return _BNTrainFn.apply(x, stats, gp, bp, running_mean, running_var, float(momentum), float(eps), int(relu),
                                residual)
    coef = torch.empty(2 * c, dtype=torch.float32, device=x.device)
    rm = pad_vec(running_mean, c, 0.0).contiguous()
    rv = pad_vec(running_var, c, 1.0).contiguous()
    call("npp_bn_eval_coef", fptr(gp.detach().contiguous() if gp is not None else None),
         fptr(bp.detach().contiguous() if bp is not None else None), fptr(rm), fptr(rv), f32(eps), fptr(coef[:c]),
         fptr(coef[c:]), i32(c), stream())
    y = _AffineFn.apply(x, coef[:c], coef[c:], int(relu))
    if residual is not None:
        y = add(y, residual)
    return y



# ------------------------------------------------------------------------------------------------
# fused cell node: y = f_a(a) [+ f_b(b)] with f = pending BatchNorm or identity; outputs raw and/or relu
# ------------------------------------------------------------------------------------------------
class Pending:
    """The input of a BatchNorm2d whose normalisation has not been applied yet: `y` (internal tensor, usually a
    conv output), `stats` (per-channel sum / sum of squares from the conv epilogue, or None) and the BatchNorm
    module `bn`.  Consumers fold the normalisation into their own pass (node()); anything else calls
    to_internal() / finish(), which applies it once and caches the result."""
    __slots__ = ("y", "stats", "bn", "raw", "relu", "fan", "spare")

    def __init__(self, y, stats, bn):
        self.y, self.stats, self.bn = y, stats, bn
        self.raw = self.relu = None
        self.fan, self.spare = 1, []      # consumers of relu(.) announced by set_fanout / handles not handed out yet

    @property
    def shape(self):
        return self.y.shape


def set_fanout(p, n):
    """Announces that relu(p) of a Pending BatchNorm output will be read by n consumers: each gets its own handle, so
    their gradients are summed inside the node's backward kernel instead of by autograd add kernels."""
    if isinstance(p, Pending) and p.relu is None:
        p.fan = max(1, int(n))
    return p


def materialize(x):
    """Computes relu(bn(.)) of a Pending BatchNorm output NOW (on the current stream) without handing a consumer
    handle out; relu(x) then only distributes handles.  Used before a TaskStreams fork so that the tensor exists — and
    is marked for the other stream — before consumers on both streams ask for it."""
    if isinstance(x, Pending) and x.relu is None:
        if x.fan > 1:    # several consumers announced (set_fanout): one handle each, gradients summed in the node
            _, rels = node(x, None, want_raw=False, want_relu=True, fan=(1, x.fan))
            x.relu, x.spare = rels[-1], list(rels)       # the canonical handle doubles as the last one handed out
        else:
            _, x.relu = node(x, None, want_raw=False, want_relu=True)
    return x


def finish(p):
    """Normalised (raw) value of a Pending BatchNorm output."""
    if not isinstance(p, Pending):
        return p
    if p.raw is None:
        p.raw, _ = node(p, None, want_raw=True, want_relu=False)
    return p.raw


def alias(buf, c_off, c):
    """Channel slice [c_off, c_off+c) of an internal tensor as a fresh tensor on the same storage (not an
    autograd view: the node kernels write it, autograd sees an ordinary output)."""
    n, _, h, w = buf.shape
    sn, sc, sh, sw = buf.stride()
    return torch.empty(0, dtype=buf.dtype, device=buf.device).set_(
        buf.untyped_storage(), buf.storage_offset() + c_off, (n, c, h, w), (sn, sc, sh, sw))


class _BNSide:
    """Per-input BatchNorm bookkeeping of one node call (forward coefficients, saved statistics)."""
    __slots__ = ("bn", "stats", "coef", "c")


def _bn_training(bn):
    return bn is not None and (bn.training or not bn.track_running_stats)


def _bn_local_stats(y, stats):
    """Per-channel (sum, sum of squares) of y: the conv epilogue's vector when it exists, else one npp_bn_stats pass."""
    if stats is None or stats.numel() == 0:
        stats = zeros_f32(2 * y.shape[1], y.device)
        call("npp_bn_stats", ref(view(y)), fptr(stats), i32(L.dtype_code(y)), stream())
    return stats


def _bn_forward_coef(y, stats, bn, sync, defer=False, reduced=False):
    """Batch statistics -> (scale, shift, mean, invstd) as one [4C] tensor; updates the running statistics
    (nn.BatchNorm2d training semantics, momentum 0.1 / eps 1e-5 in operations.py:27).
    defer=True (training): the finalize arithmetic is left to the consuming node kernel (npp_node_fwd_bn);
    returns (coef, count, gamma, fin) with fin = the npp_bn_fin descriptor, or fin = None when the coefficients were
    computed here.  reduced=True (SyncBN): `stats` already holds the sums over all ranks (the caller exchanged the
    vectors of several BatchNorms in one message); only the sample count is scaled here."""
    n, c, h, w = y.shape
    dev = y.device
    training = bn.training or not bn.track_running_stats
    gamma = pad_vec(bn.weight.detach() if bn.weight is not None else None, c, 1.0)
    beta = pad_vec(bn.bias.detach() if bn.bias is not None else None, c, 0.0)
    coef = torch.empty(4 * c, dtype=torch.float32, device=dev)
    if not training:
        rm = pad_vec(bn.running_mean, c, 0.0).contiguous()
        rv = pad_vec(bn.running_var, c, 1.0).contiguous()
        call("npp_bn_eval_coef", fptr(gamma.contiguous() if gamma is not None else None),
             fptr(beta.contiguous() if beta is not None else None), fptr(rm), fptr(rv), f32(bn.eps), fptr(coef[:c]),
             fptr(coef[c:2 * c]), i32(c), stream())
        if torch.is_grad_enabled():  # backward through running statistics: xhat = (x - rm) * rsqrt(rv + eps)
            coef[2 * c:3 * c] = rm
            coef[3 * c:] = torch.rsqrt(rv + bn.eps)
        return (coef, float(n * h * w), gamma, None) if defer else (coef, float(n * h * w), gamma)
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        if _state["defer_bn_counters"] is not None:
            _state["defer_bn_counters"].append(bn.num_batches_tracked)  # one foreach add per step (engine.TrainStep)
        else:
            bn.num_batches_tracked.add_(1)
    stats = _bn_local_stats(y, stats)
    count = float(n * h * w)
    if sync:
        count *= _sync_world() if reduced else _allreduce_sum(stats)
    c_run = bn.running_mean.numel() if bn.running_mean is not None else 0
    if defer:
        fin = L.BnFin(fptr(stats).value, fptr(gamma).value, fptr(beta).value, fptr(bn.running_mean).value,
                      fptr(bn.running_var).value, fptr(coef).value, float(bn.momentum), float(bn.eps), int(c_run))
        fin._keep = (stats, gamma, beta, coef)
        return coef, count, gamma, fin
    call("npp_bn_finalize", fptr(stats), f64(count), fptr(gamma), fptr(beta), fptr(bn.running_mean),
         fptr(bn.running_var), f32(bn.momentum), f32(bn.eps), fptr(coef[:c]), fptr(coef[c:2 * c]), fptr(coef[2 * c:3 * c]),
         fptr(coef[3 * c:]), i32(c), i32(c_run), stream())
    return (coef, count, gamma, None) if defer else (coef, count, gamma)


_NODE_STRIPES = max(1, min(64, int(os.environ.get("NPP_NODE_STRIPES", "8"))))


class _NodeFn(Function):
    """y = f_a(a) [+ f_b(b)]; f = BatchNorm(batch statistics) when the side is Pending, identity otherwise.
    Returns (raw, relu) — either may be None.  csrc/node.cu."""

    @staticmethod
    def forward(ctx, a, ga, ba, b, gb, bb, cfg):
        bn_a, st_a, bn_b, st_b, want_raw, want_relu, out_raw, out_relu, pslots, fan = cfg
        ctx.set_materialize_grads(False)
        ctx.pslots = pslots
        n, c, h, w = a.shape
        code = L.dtype_code(a)
        sync = _sync_group() is not None
        ca = cb = None
        count = float(n * h * w)
        gam_a = gam_b = None
        fin_a = fin_b = None
        defer = _state.get("node_fused_finalize", True)
        reduced = False
        if sync:
            # SyncBN: the statistic vectors of BOTH BatchNorms of the node travel in one exchange
            vecs = []
            if _bn_training(bn_a):
                st_a = _bn_local_stats(a, st_a)
                vecs.append(st_a)
            if _bn_training(bn_b):
                st_b = _bn_local_stats(b, st_b)
                vecs.append(st_b)
            if vecs:
                _allreduce_sum(*vecs)
                reduced = True
        if bn_a is not None:
            ca, count, gam_a, fin_a = _bn_forward_coef(a, st_a, bn_a, sync, defer=defer, reduced=reduced)
        if bn_b is not None:
            cb, count, gam_b, fin_b = _bn_forward_coef(b, st_b, bn_b, sync, defer=defer, reduced=reduced)
        raw = (out_raw if out_raw is not None else empty_internal(n, c, h, w, a.dtype, a.device)) if want_raw else None
        rel = (out_relu if out_relu is not None else empty_internal(n, c, h, w, a.dtype, a.device)) if want_relu else None
        if fin_a is not None or fin_b is not None:
            # BatchNorm finalize (batch sums -> scale / shift / mean / invstd, running statistics) inside the node kernel
            call("npp_node_fwd_bn", ref(view(a)), ref(fin_a) if fin_a is not None else NULL,
                 fptr(ca[:c]) if (ca is not None and fin_a is None) else NULL,
                 fptr(ca[c:2 * c]) if (ca is not None and fin_a is None) else NULL,
                 ref(view(b)) if b is not None else NULL, ref(fin_b) if fin_b is not None else NULL,
                 fptr(cb[:c]) if (cb is not None and fin_b is None) else NULL,
                 fptr(cb[c:2 * c]) if (cb is not None and fin_b is None) else NULL,
                 ref(view(raw)) if raw is not None else NULL, ref(view(rel)) if rel is not None else NULL, f64(count),
                 i32(code), stream(), keep=(fin_a, fin_b))
        else:
            call("npp_node_fwd", ref(view(a)), fptr(ca[:c]) if ca is not None else NULL,
                 fptr(ca[c:2 * c]) if ca is not None else NULL, ref(view(b)) if b is not None else NULL,
                 fptr(cb[:c]) if cb is not None else NULL, fptr(cb[c:2 * c]) if cb is not None else NULL,
                 ref(view(raw)) if raw is not None else NULL, ref(view(rel)) if rel is not None else NULL, i32(code),
                 stream())
        eval_a = bn_a is not None and not (bn_a.training or not bn_a.track_running_stats)
        eval_b = bn_b is not None and not (bn_b.training or not bn_b.track_running_stats)
        ctx.flags = (bn_a is not None, bn_b is not None, b is not None, count, sync, eval_a, eval_b)
        ctx.save_for_backward(a if bn_a is not None else None, ca, gam_a, b if bn_b is not None else None, cb, gam_b,
                              rel)
        # One handle per CONSUMER of either output (further primitives of the cell, the concat route of
        # functional.assemble): the consumers' gradients then reach backward() separately and are summed inside the
        # node's own backward kernel (npp_node_bwd_reduce3) instead of by autograd's accumulation kernels.
        n_raw = (max(1, fan[0]) if raw is not None else 0)
        n_rel = (max(1, fan[1]) if rel is not None else 0)
        ctx.fan = (n_raw, n_rel)
        outs = []
        if raw is not None:
            outs += [raw] + [alias(raw, 0, c) for _ in range(n_raw - 1)]
        if rel is not None:
            outs += [rel] + [alias(rel, 0, c) for _ in range(n_rel - 1)]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        a, ca, gam_a, b, cb, gam_b, rel = ctx.saved_tensors
        has_a, has_b, two, count, sync, eval_a, eval_b = ctx.flags
        n_raw, n_rel = ctx.fan
        raws = [g for g in grads[:n_raw] if g is not None]
        rels = [g for g in grads[n_raw:n_raw + n_rel] if g is not None]
        if not raws and not rels:
            return None, None, None, None, None, None, None
        like = rel if rel is not None else (a if a is not None else (raws[0] if raws else rels[0]))
        raws = [as_internal_grad(g, like) for g in raws]
        rels = [as_internal_grad(g, like) for g in rels]
        g_raw = raws[0] if raws else None
        g_relu = rels[0] if rels else None
        # up to two more gradients ride along as typed extra slots of the reduce kernel; any beyond that are folded
        # into one of the slots first (rare: an output with more than three consumers)
        extras = [(g, 0) for g in raws[1:]] + [(g, 1) for g in rels[1:]]
        while len(extras) > 2:
            g_last, k_last = extras.pop()
            j = next((i for i, (_, k) in enumerate(extras) if k == k_last), None)
            if j is not None:
                extras[j] = (add(extras[j][0], g_last), k_last)
            elif k_last:
                g_relu = add(g_relu, g_last)
            else:
                g_raw = add(g_raw, g_last)
        # the typed entry point (npp_node_bwd_reduce2: slot 0 raw, slot 1 relu) serves the striped / legacy paths
        typed = all(k == i for i, (_, k) in enumerate(extras)) if len(extras) == 2 else (len(extras) == 1)
        g_raw2 = g_relu2 = None
        if extras and typed and len(extras) == 1:
            g_raw2, g_relu2 = (None, extras[0][0]) if extras[0][1] else (extras[0][0], None)
        elif extras and typed:
            g_raw2, g_relu2 = extras[0][0], extras[1][0]
        ref_t = g_raw if g_raw is not None else g_relu
        n, c, h, w = ref_t.shape
        code = L.dtype_code(ref_t)
        dev = ref_t.device
        g = g_raw
        need_bn = has_a or has_b
        combine = g_relu is not None or bool(extras)
        if combine or need_bn:
            if combine:
                g = empty_internal(n, c, h, w, ref_t.dtype, dev)
            nq = 2 * (int(has_a) + int(has_b))
            parts = sums = None
            sl_a, sl_b = ctx.pslots
            # npp_node_bwd_reduce_atomic (per-block sums added with atomics, no partials buffer / fold kernel) was
            # measured SLOWER than partials + fold (15.7 vs 13.8 ms per step: ~600 blocks hammer the same 4C
            # addresses), so the deterministic path stays; the switch is kept for experiments
            atomic = (need_bn and _arena.active and _state.get("node_reduce_atomics", False) and not extras)
            # striped path (NPP_NODE_STRIPED=1): reduce blocks add into striped copies of the totals, the apply
            # kernel folds them and writes d beta / d gamma into the flat gradient buffer — no partials buffer, no
            # fold kernel (326 launches per step), but measured slower end to end (see _state); SyncBN needs the
            # totals between the two kernels and always takes the partials path.
            striped = (need_bn and not sync and not (eval_a or eval_b) and not atomic
                       and _state.get("node_striped", True) and (not extras or typed))
            if need_bn and not atomic and not striped:
                nblk = L.lib().npp_node_bwd_blocks(i32(n), i32(h), i32(w), i32(c), i32(code))
                parts = torch.empty(nblk * nq * c, dtype=torch.float32, device=dev)
            common = (ref(view(g_raw)) if g_raw is not None else NULL,
                      ref(view(g_relu)) if g_relu is not None else NULL, ref(view(rel)) if g_relu is not None else NULL,
                      ref(view(a)) if has_a else NULL, fptr(ca[2 * c:3 * c]) if has_a else NULL,
                      fptr(ca[3 * c:]) if has_a else NULL, ref(view(b)) if has_b else NULL,
                      fptr(cb[2 * c:3 * c]) if has_b else NULL, fptr(cb[3 * c:]) if has_b else NULL,
                      ref(view(g)) if combine else NULL)
            common2 = (common[0], ref(view(g_raw2)) if g_raw2 is not None else NULL, common[1],
                       ref(view(g_relu2)) if g_relu2 is not None else NULL) + common[2:]
            if striped:
                sums = zeros_f32(_NODE_STRIPES * nq * c, dev)
                call("npp_node_bwd_reduce2", *common2, NULL, fptr(sums), i32(_NODE_STRIPES), i32(code), stream())
            elif extras:
                ex = extras + [(None, 0)] * (2 - len(extras))
                call("npp_node_bwd_reduce3", common[0], common[1], common[2],
                     ref(view(ex[0][0])), i32(ex[0][1]), ref(view(ex[1][0])) if ex[1][0] is not None else NULL, i32(ex[1][1]),
                     *common[3:], fptr(parts), i32(code), stream())
            elif atomic:
                sums = zeros_f32(nq * c, dev)
                segs = []
                if has_a:
                    segs += list(sl_a) if sl_a is not None else [None, None]
                if has_b:
                    segs += list(sl_b) if sl_b is not None else [None, None]
                accp = (ctypes.c_void_p * len(segs))(*[t.data_ptr() if t is not None else None for t in segs])
                accv = (ctypes.c_int * len(segs))(*[min(t.numel(), c) if t is not None else 0 for t in segs])
                call("npp_node_bwd_reduce_atomic", *common, fptr(sums), accp, accv, i32(code), stream())
            else:
                call("npp_node_bwd_reduce", *common, fptr(parts), i32(code), stream())
        da = db = dga = dba = dgb = dbb = None
        if need_bn and striped:
            sl_a, sl_b = ctx.pslots
            segs = []
            if has_a:
                segs += list(sl_a) if sl_a is not None else [None, None]
            if has_b:
                segs += list(sl_b) if sl_b is not None else [None, None]
            accp = (ctypes.c_void_p * len(segs))(*[t.data_ptr() if t is not None else None for t in segs])
            accv = (ctypes.c_int * len(segs))(*[min(t.numel(), c) if t is not None else 0 for t in segs])
            if has_a:
                da = torch.empty_like(a)
            if has_b:
                db = torch.empty_like(b)
            call("npp_node_bwd_apply_striped", ref(view(g)), ref(view(a)) if has_a else NULL,
                 fptr(gam_a) if has_a else NULL, fptr(ca[2 * c:3 * c]) if has_a else NULL,
                 fptr(ca[3 * c:]) if has_a else NULL, ref(view(da)) if has_a else NULL,
                 ref(view(b)) if has_b else NULL, fptr(gam_b) if has_b else NULL,
                 fptr(cb[2 * c:3 * c]) if has_b else NULL, fptr(cb[3 * c:]) if has_b else NULL,
                 ref(view(db)) if has_b else NULL, fptr(sums), i32(_NODE_STRIPES), accp, accv, f64(count), i32(code),
                 stream(), keep=segs)
            if (has_a and sl_a is None) or (has_b and sl_b is None):   # gradients returned through autograd
                local = sums.view(_NODE_STRIPES, nq * c).sum(0)
                if has_a:
                    dba, dga = local[:c], local[c:2 * c]
                if has_b:
                    o = 2 * c * int(has_a)
                    dbb, dgb = local[o:o + c], local[o + c:o + 2 * c]
        elif need_bn:
            if sums is None:   # deterministic path: fold the per-block partials
                sums = torch.empty(nq * c, dtype=torch.float32, device=dev)
                if (has_a and sl_a is not None) or (has_b and sl_b is not None):
                    segs = []
                    if has_a:
                        segs += list(sl_a) if sl_a is not None else [None, None]
                    if has_b:
                        segs += list(sl_b) if sl_b is not None else [None, None]
                    accp = (ctypes.c_void_p * len(segs))(*[t.data_ptr() if t is not None else None for t in segs])
                    accv = (ctypes.c_int * len(segs))(*[min(t.numel(), c) if t is not None else 0 for t in segs])
                    call("npp_reduce_partials_acc", fptr(parts), i32(nblk), i32(nq * c), fptr(sums), i32(c), accp,
                         accv, stream())
                else:
                    call("npp_reduce_partials", fptr(parts), i32(nblk), i32(nq * c), fptr(sums), stream())
            local = sums
            if sync:
                returns_grads = (has_a and ctx.pslots[0] is None) or (has_b and ctx.pslots[1] is None)
                if returns_grads:
                    local = sums.clone()  # parameter gradients stay local (the gradient all-reduce averages them)
                _allreduce_sum(sums)
            sa = sums[:2 * c] if has_a else None
            sb = sums[2 * c * int(has_a):] if has_b else None
            if eval_a or eval_b:  # running statistics do not depend on the batch: d_in = gamma * invstd * g
                if local is sums:
                    local = sums.clone()
                if eval_a:
                    sa.zero_()
                if eval_b:
                    sb.zero_()
            if has_a:
                da = torch.empty_like(a)
                dba, dga = local[:c], local[c:2 * c]
            if has_b:
                db = torch.empty_like(b)
                o = 2 * c * int(has_a)
                dbb, dgb = local[o:o + c], local[o + c:o + 2 * c]
            call("npp_node_bwd_apply", ref(view(g)), ref(view(a)) if has_a else NULL, fptr(gam_a) if has_a else NULL,
                 fptr(ca[2 * c:3 * c]) if has_a else NULL, fptr(ca[3 * c:]) if has_a else NULL, fptr(sa),
                 ref(view(da)) if has_a else NULL, ref(view(b)) if has_b else NULL, fptr(gam_b) if has_b else NULL,
                 fptr(cb[2 * c:3 * c]) if has_b else NULL, fptr(cb[3 * c:]) if has_b else NULL, fptr(sb),
                 ref(view(db)) if has_b else NULL, f64(count), i32(code), stream())
        if not has_a:
            da = g
        if two and not has_b:
            db = g
        ni = ctx.needs_input_grad
        if need_bn:
            if has_a and ctx.pslots[0] is not None:
                dga = dba = None   # already accumulated into the parameters' gradient slots
            if has_b and ctx.pslots[1] is not None:
                dgb = dbb = None
        return (da if ni[0] else None, dga if ni[1] else None, dba if ni[2] else None, db if (two and ni[3]) else None,
                dgb if ni[4] else None, dbb if ni[5] else None, None)


def _side(x):
    if isinstance(x, Pending):
        return x.y, x.bn, x.stats
    return check_raw(x, "node"), None, None


def node(a, b=None, want_raw=True, want_relu=False, out_raw=None, out_relu=None, want_cat=False, fan=None):
    """One fused pass for a cell node (model_augment.py:48-62): a and b are internal tensors or Pending BatchNorm
    outputs; returns (raw, relu) — either None when not wanted — with `raw._npp_relu = relu` when both exist.
    out_raw / out_relu: preallocated destinations (channel slices of a concat buffer, see alias()).
    want_cat: also return second handles (raw_c, relu_c) on the same storage for the concat route — the gradient that
    comes down through the concat buffer then stays separate from the in-cell consumers' gradients until the node's
    own backward kernel adds them (no strided autograd add).
    fan=(n_raw, n_relu): return ([raw handles], [relu handles]) with one handle per consumer of either output (the
    general form of want_cat): every consumer's gradient reaches the node's backward kernel separately and is summed
    there, not by autograd's accumulation kernels."""
    if b is not None and not isinstance(a, Pending) and isinstance(b, Pending):
        a, b = b, a  # keep a BatchNorm side first (the kernels take either layout; this just normalises)
    ya, bn_a, st_a = _side(a)
    yb, bn_b, st_b = _side(b) if b is not None else (None, None, None)
    if yb is not None and ya.shape != yb.shape:
        raise RuntimeError("node: shape mismatch %s vs %s" % (tuple(ya.shape), tuple(yb.shape)))
    if not (want_raw or want_relu):
        raise RuntimeError("node: nothing to produce")
    c = ya.shape[1]

    def affine(bn):
        if bn is None:
            return None, None
        return pad_vec(bn.weight, c, 1.0), pad_vec(bn.bias, c, 0.0)

    ga, ba = affine(bn_a)
    gb, bb = affine(bn_b)
    # direct parameter-gradient slots (d beta, d gamma) per BatchNorm side; all-or-nothing per side
    pslots = []
    for bn in (bn_a, bn_b):
        sl = None
        if bn is not None and bn.affine and (bn.training or not bn.track_running_stats):
            sl = (grad_slot(bn.bias), grad_slot(bn.weight))
            if sl[0] is None or sl[1] is None:
                sl = None
        pslots.append(sl)
    cat_req = bool(want_cat)
    split = _state.get("node_cat_grads", True) and torch.is_grad_enabled()
    if fan is not None:
        n_raw, n_rel = (max(1, int(fan[0])), max(1, int(fan[1]))) if split else (1, 1)
    else:
        n_raw = n_rel = 2 if (cat_req and split) else 1
    outs = _NodeFn.apply(ya, ga, ba, yb, gb, bb, (bn_a, st_a, bn_b, st_b, bool(want_raw), bool(want_relu), out_raw,
                                                  out_relu, pslots, (n_raw, n_rel)))
    raws = list(outs[:n_raw]) if want_raw else []
    rels = list(outs[len(raws):]) if want_relu else []
    for r in rels:
        r._npp_is_relu = True   # relu(rel) is rel (no tensor ever references itself: that would leak the graph)
    raw, rel = (raws[0] if raws else None), (rels[0] if rels else None)
    if raw is not None and rel is not None:
        raw._npp_relu = rel
    if fan is not None:
        want = (max(1, int(fan[0])), max(1, int(fan[1])))
        # without gradient splitting every consumer simply shares the one handle
        raws = (raws + raws[-1:] * want[0])[:want[0]] if raws else []
        rels = (rels + rels[-1:] * want[1])[:want[1]] if rels else []
        return raws, rels
    if cat_req:
        return raw, rel, (raws[1] if len(raws) > 1 else raw), (rels[1] if len(rels) > 1 else rel)
    return raw, rel


def state_handle(raw, rel):
    """What a cell passes on as `the state`: the raw tensor (with relu attached when it exists), or — when the
    producer was told that every consumer starts with nn.ReLU — relu(state) flagged so that any raw use raises."""
    if raw is not None:
        return raw
    h = rel.detach()          # a distinct handle on the same storage; its only legitimate use is relu(h) -> rel
    h._npp_relu = rel
    h._npp_relu_only = True
    return h


class _AssembleFn(Function):
    """The concat buffer whose channel slices the node kernels have already written (torch.cat of
    model_augment.py:62 without a copy); backward hands the gradient out as channel slices."""

    @staticmethod
    def forward(ctx, holder, *slices):
        ctx.cs = [t.shape[1] for t in slices]
        return holder[0]

    @staticmethod
    def backward(ctx, dy):
        outs, off = [None], 0
        for c in ctx.cs:
            outs.append(dy[:, off:off + c])
            off += c
        return tuple(outs)


def assemble(buf, slices):
    """A handle on the concat buffer `buf` whose channel slices are `slices` (a fresh tensor on the same storage per
    call: a cell output with several consumers is assembled once per consumer, each from its own slice handles)."""
    return _AssembleFn.apply([alias(buf, 0, buf.shape[1])], *slices)

# ------------------------------------------------------------------------------------------------
# add / concat
# ------------------------------------------------------------------------------------------------
class _AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        y = empty_internal(a.shape[0], a.shape[1], a.shape[2], a.shape[3], a.dtype, a.device)
        call("npp_add", ref(view(a)), ref(view(b)), ref(view(y)), i32(L.dtype_code(a)), stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    """a + b where either side may still be a Pending BatchNorm output (normalised inside the same pass)."""
    if isinstance(a, Pending) or isinstance(b, Pending):
        return node(a, b)[0]
    check_raw(a, "add")
    check_raw(b, "add")
    if a.shape != b.shape:
        raise RuntimeError("add: shape mismatch %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    return _AddFn.apply(a, b)


class _CatFn(Function):
    @staticmethod
    def forward(ctx, *ts):
        n, _, h, w = ts[0].shape
        cs = [t.shape[1] for t in ts]
        y = empty_internal(n, sum(cs), h, w, ts[0].dtype, ts[0].device)
        off = 0
        code = L.dtype_code(y)
        for t, c in zip(ts, cs):
            call("npp_add", ref(view(t)), NULL, ref(view(y[:, off:off + c])), i32(code), stream())
            off += c
        ctx.cs = cs
        return y

    @staticmethod
    def backward(ctx, dy):
        outs, off = [], 0
        for c in ctx.cs:
            outs.append(dy[:, off:off + c])
            off += c
        return tuple(outs)


def cat(ts):
    """torch.cat(dim=1) of internal tensors (model_augment.py:62); backward hands out channel slices."""
    ts = [check_raw(to_internal(t), "cat") for t in ts]
    if len(ts) == 1:
        return ts[0]
    return _CatFn.apply(*ts)


def cat_relu(ts):
    """relu(torch.cat(ts, dim=1)) written directly (one pass per input, no intermediate concat); returns a state
    handle that only ReLU-first consumers may read."""
    ts = [check_raw(to_internal(t), "cat") for t in ts]
    n, _, h, w = ts[0].shape
    buf = empty_internal(n, sum(t.shape[1] for t in ts), h, w, ts[0].dtype, ts[0].device)
    outs, off = [], 0
    for t in ts:
        c = t.shape[1]
        outs.append(node(t, None, want_raw=False, want_relu=True, out_relu=alias(buf, off, c))[1])
        off += c
    return state_handle(None, assemble(buf, outs))


# ------------------------------------------------------------------------------------------------
# pooling
# ------------------------------------------------------------------------------------------------
class _MaxPool3Fn(Function):
    @staticmethod
    def forward(ctx, x, stride):
        n, c, h, w = x.shape
        ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        y = empty_internal(n, c, ho, wo, x.dtype, x.device)
        idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if ctx.needs_input_grad[0] else None
        call("npp_maxpool3x3_fwd", ref(view(x)), ref(view(y)), fptr(idx), i32(stride), i32(L.dtype_code(x)), stream())
        ctx.stride, ctx.shape = stride, tuple(x.shape)
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        dy = as_internal_grad(dy, dy)
        n, c, h, w = ctx.shape
        dx = empty_internal(n, c, h, w, dy.dtype, dy.device)
        call("npp_maxpool3x3_bwd", fptr(idx), ref(view(dy)), ref(view(dx)), i32(ctx.stride), i32(L.dtype_code(dy)),
             stream())
        return dx, None


class _AvgPool3Fn(Function):
    @staticmethod
    def forward(ctx, x, stride):
        n, c, h, w = x.shape
        ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        y = empty_internal(n, c, ho, wo, x.dtype, x.device)
        call("npp_avgpool3x3_fwd", ref(view(x)), ref(view(y)), i32(stride), i32(L.dtype_code(x)), stream())
        ctx.stride, ctx.shape = stride, x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_internal_grad(dy, dy)
        n, c, h, w = ctx.shape
        dx = empty_internal(n, c, h, w, dy.dtype, dy.device)
        call("npp_avgpool3x3_bwd", ref(view(dy)), ref(view(dx)), i32(ctx.stride), i32(L.dtype_code(dy)), stream())
        return dx, None


class _AvgPool2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        n, c, h, w = x.shape
        y = empty_internal(n, c, h // 2, w // 2, x.dtype, x.device)
        call("npp_avgpool2x2_fwd", ref(view(x)), ref(view(y)), i32(L.dtype_code(x)), stream())
        ctx.shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_internal_grad(dy, dy)
        n, c, h, w = ctx.shape
        dx = empty_internal(n, c, h, w, dy.dtype, dy.device)
        call("npp_avgpool2x2_bwd", ref(view(dy)), ref(view(dx)), i32(L.dtype_code(dy)), stream())
        return dx


def max_pool3x3(x, stride):
    return _MaxPool3Fn.apply(x, int(stride))


def avg_pool3x3(x, stride):
    return _AvgPool3Fn.apply(x, int(stride))


def avg_pool2x2(x):
    return _AvgPool2Fn.apply(x)


# ------------------------------------------------------------------------------------------------
# SE block
# ------------------------------------------------------------------------------------------------
class _SEFn(Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, slots):
        n, c, h, w = x.shape
        dev, code = x.device, L.dtype_code(x)
        ctx.slots = slots
        g = zeros_f32(n * c, dev).view(n, c)
        call("npp_gap_fwd", ref(view(x)), fptr(g), i32(code), stream())
        w1c, w2c = w1.detach().reshape(c // 2, c).contiguous(), w2.detach().reshape(c, c // 2).contiguous()
        b1c = b1.detach().contiguous() if b1 is not None else None
        b2c = b2.detach().contiguous() if b2 is not None else None
        hbuf = torch.empty((n, c // 2), dtype=torch.float32, device=dev)
        s = torch.empty((n, c), dtype=torch.float32, device=dev)
        call("npp_se_fc_fwd", fptr(g), fptr(w1c), fptr(b1c), fptr(w2c), fptr(b2c), fptr(hbuf), fptr(s), i32(n), i32(c),
             stream())
        y = empty_internal(n, c, h, w, x.dtype, dev)
        call("npp_se_scale_fwd", ref(view(x)), fptr(s), ref(view(y)), i32(code), stream())
        ctx.save_for_backward(x, g, hbuf, s, w1c, w2c)
        ctx.has_bias = (b1 is not None, b2 is not None)
        ctx.wshapes = (w1.shape, w2.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, hbuf, s, w1c, w2c = ctx.saved_tensors
        dy = as_internal_grad(dy, x)
        n, c, h, w = x.shape
        dev, code = x.device, L.dtype_code(x)
        ds = zeros_f32(n * c, dev).view(n, c)
        call("npp_se_bwd_reduce", ref(view(x)), ref(view(dy)), fptr(ds), i32(code), stream())
        direct = ctx.slots is not None
        if direct:   # se_fc_bwd accumulates (+=): straight into the four parameters' gradient slots
            dw1, db1, dw2, db2 = ctx.slots
        else:
            dw1, dw2 = torch.zeros_like(w1c), torch.zeros_like(w2c)
            db1 = zeros_f32(c // 2, dev)
            db2 = zeros_f32(c, dev)
        dg = torch.empty((n, c), dtype=torch.float32, device=dev)
        if _state.get("se_bwd2", False):   # round-2 candidate: no weight-gradient atomics (csrc/se.cu)
            scratch = torch.empty(n * (c + c // 2), dtype=torch.float32, device=dev)
            call("npp_se_fc_bwd2", fptr(g), fptr(hbuf), fptr(s), fptr(ds), fptr(w1c), fptr(w2c), fptr(dw1), fptr(db1),
                 fptr(dw2), fptr(db2), fptr(dg), fptr(scratch), i32(n), i32(c), stream())
        else:
            call("npp_se_fc_bwd", fptr(g), fptr(hbuf), fptr(s), fptr(ds), fptr(w1c), fptr(w2c), fptr(dw1), fptr(db1),
                 fptr(dw2), fptr(db2), fptr(dg), i32(n), i32(c), stream())
        dx = torch.empty_like(x)
        call("npp_se_bwd_apply", ref(view(dy)), fptr(s), fptr(dg), ref(view(dx)), i32(code), stream())
        if direct:
            return dx, None, None, None, None, None
        return (dx, dw1.reshape(ctx.wshapes[0]), db1 if ctx.has_bias[0] else None, dw2.reshape(ctx.wshapes[1]),
                db2 if ctx.has_bias[1] else None, None)


def se_scale(x, w1, b1, w2, b2):
    slots = tuple(grad_slot(p) for p in (w1, b1, w2, b2))
    return _SEFn.apply(x, w1, b1, w2, b2, slots if all(t is not None for t in slots) else None)


# ------------------------------------------------------------------------------------------------
# resampling
# ------------------------------------------------------------------------------------------------
class _ResampleFn(Function):
    @staticmethod
    def forward(ctx, x, oh, ow, mode, align, sh, sw):
        n, c, h, w = x.shape
        y = empty_internal(n, c, oh, ow, x.dtype, x.device)
        code = L.dtype_code(x)
        if mode == "bilinear":
            call("npp_bilinear_fwd", ref(view(x)), ref(view(y)), i32(align), f64(sh), f64(sw), i32(code), stream())
        else:
            call("npp_nearest_fwd", ref(view(x)), ref(view(y)), f64(sh), f64(sw), i32(code), stream())
        ctx.cfg = (x.shape, mode, align, sh, sw)
        return y

    @staticmethod
    def backward(ctx, dy):
        shape, mode, align, sh, sw = ctx.cfg
        dy = as_internal_grad(dy, dy)
        n, c, h, w = shape
        dx = empty_internal(n, c, h, w, dy.dtype, dy.device)
        code = L.dtype_code(dy)
        if (mode == "bilinear" and _state.get("bilinear_sep", True) and dy.shape[2] >= 2 * h and dy.shape[3] >= 2 * w):
            # separable two-pass form (csrc/resample.cu): reads dY once, fp32 scratch of N*Ho*Wi*C floats.  Only for
            # up-sampling by >= 2: when dY is not larger than dx the gather form has nothing to re-read and the
            # scratch round trip would be pure overhead
            tmp = torch.empty(n * dy.shape[2] * w * c, dtype=torch.float32, device=dy.device)
            call("npp_bilinear_bwd_sep", ref(view(dy)), ref(view(dx)), fptr(tmp), i32(align), f64(sh), f64(sw), i32(code),
                 stream())
        elif mode == "bilinear":
            call("npp_bilinear_bwd", ref(view(dy)), ref(view(dx)), i32(align), f64(sh), f64(sw), i32(code), stream())
        else:
            call("npp_nearest_bwd", ref(view(dy)), ref(view(dx)), f64(sh), f64(sw), i32(code), stream())
        return dx, None, None, None, None, None, None


def interpolate(x, scale_factor=None, size=None, mode="nearest", align_corners=None):
    """F.interpolate for internal tensors (bilinear / nearest), output size floor(in*scale) as ATen."""
    import math
    n, c, h, w = x.shape
    if size is not None:
        oh, ow = (size, size) if isinstance(size, int) else size
        sh = sw = 0.0
    else:
        sf = scale_factor if isinstance(scale_factor, (tuple, list)) else (scale_factor, scale_factor)
        sh, sw = float(sf[0]), float(sf[1])
        oh, ow = int(math.floor(h * sh)), int(math.floor(w * sw))
    if mode not in ("bilinear", "nearest"):
        raise RuntimeError("interpolate mode %r is not implemented" % mode)
    return _ResampleFn.apply(x, int(oh), int(ow), mode, int(bool(align_corners)), sh, sw)


# ------------------------------------------------------------------------------------------------
# Zero primitive
# ------------------------------------------------------------------------------------------------
class _ZeroFn(Function):
    @staticmethod
    def forward(ctx, x):
        y = empty_internal(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.dtype, x.device)
        call("npp_fill_zero", ref(view(y)), i32(L.dtype_code(y)), stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dx = torch.empty_like(dy)
        call("npp_fill_zero", ref(view(dx)), i32(L.dtype_code(dx)), stream())
        return dx


def zeros_like_internal(x):
    """`x * 0.` of the Zero primitive (operations.py:31-41): zero output, zero gradient."""
    return _ZeroFn.apply(x)


# ------------------------------------------------------------------------------------------------
# search supernet: weighted n-ary sums (MixedOp, beta-weighted nodes), fan-out, channel halves   (csrc/mix.cu)
# ------------------------------------------------------------------------------------------------
MIX_MAX = L.NPP_MIX_MAX


def _mix_desc(ys, interleave):
    d = L.MixDesc()
    d.k = len(ys)
    d.interleave = int(bool(interleave))
    for j, y in enumerate(ys):
        d.y[j] = view(y)
    return d


class _MixFn(Function):
    """out = sum_k w[k] * f_k(y_k), f_k = BatchNorm(batch statistics, affine=False) for Pending branches, identity
    otherwise; with `pass_` the result is interleaved with it (cat + channel_shuffle(2),
    model_search_interact.py:70-71).  Gradients: d w (the architecture weights), d pass, d y_k."""

    @staticmethod
    def forward(ctx, wts, pass_, cfg, *ys):
        bns, stats = cfg
        ctx.set_materialize_grads(False)
        k = len(ys)
        n, c, h, w = ys[0].shape
        code = L.dtype_code(ys[0])
        dev = ys[0].device
        sync = _sync_group() is not None
        interleave = pass_ is not None
        count = float(n * h * w)
        coefs = []
        stats = list(stats)
        reduced = False
        if sync:   # SyncBN: the branches' statistic vectors travel two per exchange
            vecs = []
            for j, (y, bn) in enumerate(zip(ys, bns)):
                if _bn_training(bn):
                    stats[j] = _bn_local_stats(y, stats[j])
                    vecs.append(stats[j])
            for j in range(0, len(vecs), 2):
                _allreduce_sum(*vecs[j:j + 2])
            reduced = bool(vecs)
        for y, bn, st in zip(ys, bns, stats):
            if bn is None:
                coefs.append(None)
            else:
                cf, count, _ = _bn_forward_coef(y, st, bn, sync, reduced=reduced)
                coefs.append(cf)
        d = _mix_desc(ys, interleave)
        for j, cf in enumerate(coefs):
            if cf is not None:
                d.scale[j] = cf[:c].data_ptr()
                d.shift[j] = cf[c:2 * c].data_ptr()
        out = empty_internal(n, 2 * c if interleave else c, h, w, ys[0].dtype, dev)
        wv = wts.detach().float().contiguous() if wts is not None else None
        call("npp_mix_fwd", ref(d), fptr(wv), ref(view(pass_)) if interleave else NULL, ref(view(out)), i32(code),
             stream())
        ctx.meta = (k, interleave, count, sync, [cf is not None for cf in coefs], wts is not None,
                    pass_.shape if interleave else None)
        ctx.save_for_backward(wv, *ys, *[cf for cf in coefs if cf is not None])
        return out

    @staticmethod
    def backward(ctx, g):
        k, interleave, count, sync, is_bn, has_w, pass_shape = ctx.meta
        none = (None,) * (3 + k)
        if g is None:
            return none
        saved = ctx.saved_tensors
        wv, ys, cfs = saved[0], saved[1:1 + k], list(saved[1 + k:])
        coefs = [cfs.pop(0) if b else None for b in is_bn]
        n, c, h, w = ys[0].shape
        code = L.dtype_code(ys[0])
        dev = ys[0].device
        g = as_internal_grad(g, ys[0])
        ni = ctx.needs_input_grad
        d = _mix_desc(ys, interleave)
        for j, cf in enumerate(coefs):
            if cf is not None:
                d.mean[j] = cf[2 * c:3 * c].data_ptr()
                d.invstd[j] = cf[3 * c:].data_ptr()
        nblk = L.lib().npp_node_bwd_blocks(i32(n), i32(h), i32(w), i32(c), i32(code))
        parts = torch.empty(nblk * (k + 1) * c, dtype=torch.float32, device=dev)
        call("npp_mix_bwd_reduce", ref(d), ref(view(g)), fptr(parts), i32(code), stream())
        sums = torch.empty((k + 1) * c, dtype=torch.float32, device=dev)
        call("npp_reduce_partials", fptr(parts), i32(nblk), i32((k + 1) * c), fptr(sums), stream())
        dw = None
        if has_w and ni[0]:
            dw = torch.empty(k, dtype=torch.float32, device=dev)
            call("npp_mix_dw", ref(d), fptr(sums), NULL, fptr(dw), stream())   # from the LOCAL sums (DDP averages them)
        if sync and any(is_bn):
            _allreduce_sum(sums)
        dys = []
        keep = []
        for j in range(k):
            if ni[3 + j]:
                t = torch.empty_like(ys[j])
                d.dy[j] = view(t)
                dys.append(t)
            else:
                dys.append(None)
        dpass = None
        if interleave and ni[1]:
            dpass = empty_internal(pass_shape[0], pass_shape[1], pass_shape[2], pass_shape[3], ys[0].dtype, dev)
        call("npp_mix_bwd_apply", ref(d), ref(view(g)), fptr(wv), fptr(sums), f64(count),
             ref(view(dpass)) if dpass is not None else NULL, i32(code), stream())
        del keep
        return (dw, dpass, None) + tuple(dys)


def mix(branches, weights=None, pass_=None):
    """sum_k weights[k] * branch_k  (branch: internal tensor or Pending BatchNorm output), optionally interleaved
    channel-wise with `pass_` (out[2c] = sum, out[2c+1] = pass_[c]).  weights: 1-D fp32 device tensor or None (ones)."""
    if not 1 <= len(branches) <= MIX_MAX:
        raise RuntimeError("mix: between 1 and %d branches, got %d" % (MIX_MAX, len(branches)))
    ys, bns, stats = [], [], []
    for b in branches:
        if isinstance(b, Pending):
            bn = b.bn
            if bn.affine or not (bn.training or not bn.track_running_stats):
                b = finish(b)       # affine / eval-mode BatchNorm: normalised in its own pass
        if isinstance(b, Pending):
            ys.append(b.y), bns.append(b.bn), stats.append(b.stats)
        else:
            ys.append(check_raw(to_internal(b), "mix")), bns.append(None), stats.append(None)
    for y in ys[1:]:
        if y.shape != ys[0].shape:
            raise RuntimeError("mix: shape mismatch %s vs %s" % (tuple(y.shape), tuple(ys[0].shape)))
    if weights is not None and (weights.dim() != 1 or weights.numel() != len(ys)):
        raise RuntimeError("mix: need one weight per branch")
    if pass_ is not None and tuple(pass_.shape) != tuple(ys[0].shape):
        raise RuntimeError("mix: pass-through half has shape %s, branches %s" % (tuple(pass_.shape), tuple(ys[0].shape)))
    return _MixFn.apply(weights, pass_, (bns, stats), *ys)


def sum_n(ts):
    """Plain sum of internal tensors in one pass (groups of MIX_MAX)."""
    ts = list(ts)
    while len(ts) > 1:
        ts = [mix(ts[:MIX_MAX])] + ts[MIX_MAX:]
    return ts[0]


class _SplitFanFn(Function):
    """x -> (n_lo aliases of x[:, :C/2], x[:, C/2:]) without copies (MixedOp's xtemp / xtemp2,
    model_search_interact.py:59-60; xtemp is read by every candidate primitive).  Backward: the n_lo gradients are
    summed in ONE pass straight into the low half of dx, the high half is copied next to it."""

    @staticmethod
    def forward(ctx, x, n_lo):
        ctx.set_materialize_grads(False)
        c = x.shape[1]
        ctx.shape, ctx.dtype = tuple(x.shape), x.dtype
        return tuple(alias(x, 0, c // 2) for _ in range(n_lo)) + (alias(x, c // 2, c // 2),)

    @staticmethod
    def backward(ctx, *gs):
        n, c, h, w = ctx.shape
        los = [g for g in gs[:-1] if g is not None]
        hi = gs[-1]
        if not los and hi is None:
            return None, None
        ref_t = los[0] if los else hi
        dx = empty_internal(n, c, h, w, ctx.dtype, ref_t.device)
        code = L.dtype_code(dx)
        lo_v, hi_v = alias(dx, 0, c // 2), alias(dx, c // 2, c // 2)
        if los:
            los = [as_internal_grad(g, dx) for g in los]
            while len(los) > MIX_MAX:
                los = [sum_n(los[:MIX_MAX])] + los[MIX_MAX:]
            d = _mix_desc(los, False)
            call("npp_mix_fwd", ref(d), NULL, NULL, ref(view(lo_v)), i32(code), stream())
        else:
            call("npp_fill_zero", ref(view(lo_v)), i32(code), stream())
        if hi is not None:
            call("npp_add", ref(view(as_internal_grad(hi, dx))), NULL, ref(view(hi_v)), i32(code), stream())
        else:
            call("npp_fill_zero", ref(view(hi_v)), i32(code), stream())
        return dx, None


def split_halves(x, n_lo=1):
    """([n_lo handles on x[:, :C/2]], x[:, C/2:]) as channel-slice views of an internal tensor."""
    x = check_raw(to_internal(x), "split_halves")
    if x.shape[1] % 16:
        raise RuntimeError("split_halves: %d channels cannot be halved into 16-byte vectors" % x.shape[1])
    outs = _SplitFanFn.apply(x, int(n_lo))
    return list(outs[:-1]), outs[-1]


class _FanoutFn(Function):
    """n handles on one tensor; backward sums the n gradients in one pass instead of n-1 separate adds."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.set_materialize_grads(False)
        return tuple(alias(x, 0, x.shape[1]) for _ in range(n))

    @staticmethod
    def backward(ctx, *gs):
        gs = [g for g in gs if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        gs = [as_internal_grad(g, g) for g in gs]
        return sum_n(gs), None


def fanout(x, n):
    x = to_internal(x)
    if n <= 1 or not torch.is_grad_enabled() or not x.requires_grad:
        return [x] * max(n, 1)
    outs = list(_FanoutFn.apply(x, int(n)))
    for o in outs:
        for attr in ("_npp_is_relu", "_npp_relu_only"):
            if getattr(x, attr, False):
                setattr(o, attr, True)
    return outs


class _InterleaveFn(Function):
    """cat(a, b) + channel_shuffle(groups=2): out[2c] = a[c], out[2c+1] = b[c] (model_search_interact.py:22-36)."""

    @staticmethod
    def forward(ctx, a, b):
        n, c, h, w = a.shape
        y = empty_internal(n, 2 * c, h, w, a.dtype, a.device)
        call("npp_interleave2_fwd", ref(view(a)), ref(view(b)), ref(view(y)), i32(L.dtype_code(a)), stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_internal_grad(dy, dy)
        n, c2, h, w = dy.shape
        da = empty_internal(n, c2 // 2, h, w, dy.dtype, dy.device)
        db = empty_internal(n, c2 // 2, h, w, dy.dtype, dy.device)
        call("npp_interleave2_bwd", ref(view(dy)), ref(view(da)), ref(view(db)), i32(L.dtype_code(dy)), stream())
        return da, db


def interleave2(a, b):
    a, b = check_raw(to_internal(a), "interleave2"), check_raw(to_internal(b), "interleave2")
    if a.shape != b.shape:
        raise RuntimeError("interleave2: shape mismatch %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    return _InterleaveFn.apply(a, b)
