"""Data-parallel plumbing for the one way this path shards: batch data parallelism.

One process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).  Two exchange steps exist:
  * SyncBN statistics (augment_lip_sync.py:191 nn.SyncBatchNorm.convert_sync_batchnorm): our BatchNorm2d keeps the
    raw per-channel (sum, sum of squares) and (sum dy, sum dy*xhat) vectors, so synchronising is ONE all-reduce of
    2C floats per BN call in each direction (torch's SyncBatchNorm all-gathers mean/invstd/count instead);
  * gradient all-reduce (DistributedDataParallel, augment_lip_sync.py:207): engine.TrainStep reduces one flat buffer.
"""
import torch.distributed as dist

from . import functional as F_


def enable_sync_bn(group=True):
    """Makes every npp_b200.nn.BatchNorm2d in training mode use cross-rank batch statistics.
    `group`: True for the default process group, a ProcessGroup, or None/False to disable."""
    F_._state["sync_bn"] = group if group else None


def convert_sync_batchnorm(model, process_group=None):
    """Drop-in for nn.SyncBatchNorm.convert_sync_batchnorm(model): our BatchNorm2d layers are not torch
    _BatchNorm instances (torch's converter leaves them alone), synchronisation is a global switch."""
    enable_sync_bn(process_group if process_group is not None else True)
    return model


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
