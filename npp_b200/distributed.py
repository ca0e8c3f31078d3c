"""Data-parallel plumbing for the one way this path shards: batch data parallelism.

One process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the control plane and the gradient
all-reduce.  Two exchange steps exist:
  * SyncBN statistics (augment_lip_sync.py:191 nn.SyncBatchNorm.convert_sync_batchnorm): our BatchNorm2d keeps the
    raw per-channel (sum, sum of squares) and (sum dy, sum dy*xhat) vectors, so synchronising is ONE all-reduce of
    2C..4C floats per cell node in each direction (torch's SyncBatchNorm all-gathers mean/invstd/count per layer).
    These ~860 latency-bound messages per step do not go through NCCL (12-27 us each): `PeerComm` maps every rank's
    communication buffer into every peer (CUDA IPC over NVLink) and libnpp_b200's one-shot all-reduce kernel
    (csrc/peer.cu) reads the peers' staging slots directly;
  * gradient all-reduce (DistributedDataParallel, augment_lip_sync.py:207): engine.TrainStep reduces one flat buffer.
"""
import ctypes
import os
import warnings

import torch
import torch.distributed as dist

from . import _lib as L
from . import functional as F_

NPP_PEER_MAX_RANKS = 8
NPP_PEER_MAX_FLOATS = 8192


class _PeerCommStruct(ctypes.Structure):
    """Mirror of `npp_peer_comm` (include/npp_b200.h)."""
    _fields_ = [("bufs", ctypes.c_void_p * NPP_PEER_MAX_RANKS), ("rank", ctypes.c_int32), ("world", ctypes.c_int32),
                ("timeout_ms", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class PeerComm:
    """Peer-memory communicator for the SyncBN statistic vectors: one communication buffer per rank, mapped into
    every peer of `group` with CUDA IPC.  Every rank must call allreduce() in the same order (they do: replicas of
    one program).  Needs one process per GPU on one node, peer access between the GPUs and world size <= 8."""

    def __init__(self, group=None, timeout_ms=30000):
        lib = L.lib()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > NPP_PEER_MAX_RANKS:
            raise RuntimeError("PeerComm supports up to %d ranks, got %d" % (NPP_PEER_MAX_RANKS, self.world))
        self._own = ctypes.c_void_p(0)
        L.check(lib.npp_peer_alloc(ctypes.byref(self._own)), "npp_peer_alloc")
        handle = (ctypes.c_ubyte * 64)()
        L.check(lib.npp_peer_export(self._own, handle), "npp_peer_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle), os.getpid()), group=group)
        self.c = _PeerCommStruct()
        self.c.rank, self.c.world, self.c.timeout_ms = self.rank, self.world, int(timeout_ms)
        self._opened = []
        failure = None
        try:
            for p, (h, pid) in enumerate(handles):
                if p == self.rank:
                    self.c.bufs[p] = self._own.value
                    continue
                if pid == os.getpid():
                    raise RuntimeError("PeerComm needs one process per rank (CUDA IPC cannot map a buffer of its own "
                                       "process)")
                ptr = ctypes.c_void_p(0)
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                L.check(lib.npp_peer_open(buf, ctypes.byref(ptr)), "npp_peer_open")
                self._opened.append(ptr)
                self.c.bufs[p] = ptr.value
        except Exception as exc:
            failure = exc
        torch.cuda.synchronize()
        # every buffer is zeroed and mapped on EVERY rank before the first exchange — or nobody uses the transport
        ok = torch.tensor([0 if failure is not None else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            for ptr in self._opened:
                lib.npp_peer_close(ptr)
            lib.npp_peer_free(self._own)
            self._opened, self._own = [], None
            raise RuntimeError("peer mapping failed on at least one rank (this rank: %r)" % (failure,))

    def allreduce(self, t, t2=None):
        """In-place SUM over ranks of one or two fp32 device vectors (one kernel, one NVLink round trip)."""
        for v in (t, t2):
            if v is not None and not (v.is_cuda and v.dtype == torch.float32 and v.is_contiguous() and v.numel() % 4 == 0):
                raise RuntimeError("PeerComm.allreduce takes contiguous fp32 CUDA vectors with a multiple of 4 elements")
        n2 = t2.numel() if t2 is not None else 0
        if t.numel() + n2 > NPP_PEER_MAX_FLOATS:
            if t2 is not None:
                self.allreduce(t)
                self.allreduce(t2)
                return
            flat = t.view(-1)
            for off in range(0, flat.numel(), NPP_PEER_MAX_FLOATS):
                self.allreduce(flat[off:off + NPP_PEER_MAX_FLOATS])
            return
        L.call("npp_peer_allreduce", ctypes.byref(self.c), L.fptr(t), L.fptr(t), L.i32(t.numel()), L.fptr(t2), L.fptr(t2),
               L.i32(n2), L.stream())

    def status(self):
        """(exchanges issued, error) — error != 0 is the sequence number of an exchange that timed out.  Synchronises."""
        seq, err = ctypes.c_uint(0), ctypes.c_uint(0)
        L.check(L.lib().npp_peer_status(ctypes.byref(self.c), ctypes.byref(seq), ctypes.byref(err)), "npp_peer_status")
        return int(seq.value), int(err.value)

    def check(self):
        seq, err = self.status()
        if err:
            raise RuntimeError("SyncBN peer exchange %d timed out on rank %d (a peer stopped or the ranks issued "
                               "different exchange sequences)" % (err, self.rank))
        return seq

    def close(self):
        """Unmaps the peers' buffers and frees the local one.  Collective: call on every rank, after the last exchange
        has finished (CUDA graphs that captured exchanges must not be replayed afterwards)."""
        if self._own is None:
            return
        torch.cuda.synchronize()
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        lib = L.lib()
        for ptr in self._opened:
            lib.npp_peer_close(ptr)
        lib.npp_peer_free(self._own)
        self._opened, self._own = [], None


def enable_sync_bn(group=True, peer=None):
    """Makes every npp_b200.nn.BatchNorm2d in training mode use cross-rank batch statistics.
    `group`: True for the default process group, a ProcessGroup, or None/False to disable.
    `peer`: exchange the statistic vectors over NVLink peer memory (PeerComm) instead of NCCL all-reduces; default =
    whenever the process group is NCCL on CUDA with 2..8 ranks (NPP_SYNCBN_PEER=0 forces NCCL)."""
    for key in ("peer_comm", "peer_comm_side"):
        old = F_._state.get(key)
        if old is not None:
            F_._state[key] = None
            old.close()
    for old in (F_._state.pop("peer_comms", None) or {}).values():
        old.close()
    F_._state["sync_bn"] = group if group else None
    if not group or not (dist.is_available() and dist.is_initialized()):
        return
    pg = None if group is True else group
    if peer is None:
        peer = (os.environ.get("NPP_SYNCBN_PEER", "1") != "0" and torch.cuda.is_available()
                and dist.get_backend(pg) == "nccl" and 2 <= dist.get_world_size(pg) <= NPP_PEER_MAX_RANKS)
    if peer:
        try:
            F_._state["peer_comm"] = PeerComm(pg)
            if F_._state.get("two_streams"):     # the second task stream issues its own sequence of exchanges
                F_._state["peer_comm_side"] = PeerComm(pg)
                # ... and so does every worker stream of functional.parallel_branches (branch j of a node always runs on
                # worker j % n of its task stream's pool, the same on every rank): one communicator per stream role
                comms = {}
                for pool in ("main", "side"):
                    for i in range(F_._N_WORKERS if F_._N_WORKERS >= 2 else 0):
                        comms["worker:%s:%d" % (pool, i)] = PeerComm(pg)
                F_._state["peer_comms"] = comms
        except Exception as exc:   # no peer access / IPC unavailable: the NCCL transport still works
            warnings.warn("SyncBN peer-memory exchange unavailable (%r): using NCCL all-reduces" % (exc,))
            F_._state["peer_comm"] = F_._state["peer_comm_side"] = None
            F_._state["peer_comms"] = None


def peer_comm():
    return F_._state.get("peer_comm")


def convert_sync_batchnorm(model, process_group=None):
    """Drop-in for nn.SyncBatchNorm.convert_sync_batchnorm(model): our BatchNorm2d layers are not torch
    _BatchNorm instances (torch's converter leaves them alone), synchronisation is a global switch."""
    enable_sync_bn(process_group if process_group is not None else True)
    return model


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
