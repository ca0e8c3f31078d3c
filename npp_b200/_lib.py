"""ctypes binding of libnpp_b200.so (the C ABI declared in include/npp_b200.h).

The product path has no fallback: if the shared library is missing or a kernel returns an error
code, a RuntimeError is raised.  PyTorch is used for device memory (caching allocator), the
current CUDA stream and torch.distributed only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnpp_b200.so")

NPP_F32 = 0
NPP_BF16 = 1

_ERRORS = {-1: "NPP_E_INVALID (bad argument / misaligned view)", -2: "NPP_E_UNSUPPORTED (shape or dtype not implemented)",
           -3: "NPP_E_CUDA", -4: "NPP_E_NODRIVER (cuTensorMapEncodeTiled unavailable)"}


class View4(ctypes.Structure):
    """Mirror of `npp_view4` (include/npp_b200.h)."""
    _fields_ = [("ptr", ctypes.c_void_p), ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("c", ctypes.c_int32), ("sn", ctypes.c_int64), ("sh", ctypes.c_int64), ("sw", ctypes.c_int64)]


NPP_MIX_MAX = 8


class MixDesc(ctypes.Structure):
    """Mirror of `npp_mix_desc` (include/npp_b200.h)."""
    _fields_ = [("k", ctypes.c_int32), ("interleave", ctypes.c_int32), ("y", View4 * NPP_MIX_MAX),
                ("scale", ctypes.c_void_p * NPP_MIX_MAX), ("shift", ctypes.c_void_p * NPP_MIX_MAX),
                ("mean", ctypes.c_void_p * NPP_MIX_MAX), ("invstd", ctypes.c_void_p * NPP_MIX_MAX),
                ("gamma", ctypes.c_void_p * NPP_MIX_MAX), ("dy", View4 * NPP_MIX_MAX)]


class BnFin(ctypes.Structure):
    """Mirror of `npp_bn_fin` (include/npp_b200.h)."""
    _fields_ = [("stats", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
                ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p), ("coef", ctypes.c_void_p),
                ("momentum", ctypes.c_float), ("eps", ctypes.c_float), ("c_run", ctypes.c_int32)]


_lib = None


def lib():
    """Loads the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libnpp_b200.so not found at %s — build it with `python -m npp_b200.build` "
                "(there is no CPU / cuDNN fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.npp_last_error.restype = ctypes.c_char_p
        _lib.npp_version.restype = ctypes.c_char_p
        _lib.npp_launch_count.restype = ctypes.c_longlong
    return _lib


def dtype_code(t):
    if t.dtype == torch.bfloat16:
        return NPP_BF16
    if t.dtype == torch.float32:
        return NPP_F32
    raise RuntimeError("npp_b200 kernels take bf16 or fp32 activations, got %s" % t.dtype)


_pending = []        # while tracing: tensors behind the views / pointers of the ABI call being assembled
_pending_bytes = [0]  # ... and the bytes its NHWC views span (algorithmic traffic: every view once)


def view(t):
    """npp_view4 over a 4-D tensor with logical shape [N, C, H, W] and channel stride 1 (channels_last)."""
    if _trace is not None:
        _pending.append(t)
        _pending_bytes[0] += t.numel() * t.element_size()
    if t.dim() != 4:
        raise RuntimeError("expected a 4-D NCHW-shaped tensor, got %s" % (tuple(t.shape),))
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    if c > 1 and sc != 1:
        raise RuntimeError("expected a channels_last tensor (channel stride 1), got strides %s" % (t.stride(),))
    if h == 1:
        sh = w * sw
    if n == 1:
        sn = h * sh
    return View4(t.data_ptr(), n, h, w, c, sn, sh, sw)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def fptr(t):
    """Raw pointer of a dense tensor (or NULL for None)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_contiguous():
        raise RuntimeError("expected a contiguous tensor")
    if _trace is not None:
        _pending.append(t)
    return ctypes.c_void_p(t.data_ptr())


def check(rc, name):
    if rc != 0:
        msg = _ERRORS.get(rc, "error %d" % rc)
        detail = lib().npp_last_error().decode() if rc == -3 else ""
        raise RuntimeError("%s failed: %s %s" % (name, msg, detail))


_prof = None   # when profiling: list of (name, start_event, end_event, work)
_trace = None  # when tracing: list of (name, args, work, keep) of the calls that passed keep=


def call(name, *args, work=None, keep=None):
    """Invokes a C-ABI entry point.  `work` = (algorithmic flops, algorithmic bytes) of this call and `keep` = the
    tensors behind its pointer arguments; both are only used by bench.py's live roofline measurement (event
    profiler / call trace that is replayed back-to-back)."""
    fn = getattr(lib(), name)
    if _trace is not None:
        w = work or (0.0, float(_pending_bytes[0]))
        _trace.append((name, args, w, (tuple(_pending), keep)))
        del _pending[:]
        _pending_bytes[0] = 0
    if _prof is None:
        check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    _prof.append((name, e0, e1, work or (0.0, 0.0)))


def trace_begin():
    """Starts recording every ABI call with the tensors behind its arguments kept alive: (name, args, (flops,
    bytes), keep).  bytes = what the call's NHWC views span unless the caller passed work=."""
    global _trace
    _trace = []
    del _pending[:]
    _pending_bytes[0] = 0


def trace_end():
    global _trace
    t, _trace = _trace, None
    return t


def signature(trace):
    """Pointer-independent description of a recorded call sequence: [(abi name, (scalar arguments and view shapes))].
    Two runs of the same step — eager launches and the pass that a CUDA graph captures — must produce equal signatures
    (same kernels, same shapes, same order): tests/test_gpu_engine.py."""
    out = []
    for name, args, _, _ in trace:
        sig = []
        for a in args:
            obj = getattr(a, "_obj", None)
            if isinstance(obj, View4):
                sig.append(("view", obj.n, obj.h, obj.w, obj.c, obj.sw - obj.c >= 0 and obj.sw != obj.c))
            elif isinstance(a, (ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double)):
                sig.append(a.value)
        out.append((name, tuple(sig)))
    return out


def replay(trace, names):
    """Re-issues the recorded calls whose ABI name is in `names` on the current stream (same pointers, same
    shapes).  Stream arguments are re-read so the replay can be captured into a CUDA graph."""
    cur = stream()
    for name, args, _, _ in trace:
        if name in names:
            check(getattr(lib(), name)(*args[:-1], cur), name)


def profile_begin():
    """Starts bracketing every kernel-launching ABI call with CUDA events on torch's current stream."""
    global _prof
    _prof = []


def profile_end():
    """Returns {abi name: dict(calls, ms, flops, bytes)} and stops profiling."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, (fl, by) in rec:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
        d["bytes"] += by
    return out


def launch_count():
    """Number of kernels libnpp_b200 has launched in this process (bench.py's gpu_launches evidence)."""
    return int(lib().npp_launch_count())


def i32(v):
    return ctypes.c_int(int(v))


def i64(v):
    return ctypes.c_int64(int(v))


def f32(v):
    return ctypes.c_float(float(v))


def f64(v):
    return ctypes.c_double(float(v))


def ref(v):
    return ctypes.byref(v)


NULL = ctypes.c_void_p(0)
