"""Multi-GPU plumbing: one process per GPU (torchrun), patches / whole images sharded by rank, no collective on
the codec path.  The only exchange is the dataset-wide symbol statistics of get_encoded_distribution.py:113-134
(counts[q]) and cal_encoded_distribution.py:111-128 (per-position sums): one all-reduce per dataset, NCCL over
NVLink on GPUs (in place on the codec's device histogram), gloo in the CPU tests."""
from __future__ import annotations

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def world():
    """(rank, world_size) of the default process group; (0, 1) when torch.distributed is not initialised."""
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_units, rank=None, world_size=None):
    """Contiguous [begin, end) of `n_units` whole images (or patches) for this rank; sizes differ by at most one
    and the concatenation over ranks is the original order (each rank's symbol stream is a run of complete
    per-image bitstreams, encode.py:171-182)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(int(n_units), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_counts(counts, group=None, device=None):
    """Sum an integer count vector over all ranks (host array in, host array out)."""
    c = np.ascontiguousarray(np.asarray(counts, dtype=np.int64))
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return c.copy()
    t = torch.from_numpy(c.copy())
    if dist.get_backend(group) == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class _DeviceCounts:
    """uint64[256] histogram block of a Codec as a zero-copy CUDA array (tic_hist_device_ptr)."""

    def __init__(self, ptr, n=256):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def allreduce_histogram(codec, group=None):
    """Dataset-wide freq[q]: all-reduce the codec's device histogram IN PLACE over NCCL (every rank ends up with
    the global table and later hist_read() calls see it), or through the host for non-NCCL groups.
    Returns counts[q] as uint64."""
    q = codec.quan_scale
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return codec.hist_read()
    if dist.get_backend(group) == "nccl":
        codec._follow_torch_stream()
        t = torch.as_tensor(_DeviceCounts(codec.hist_device_ptr()), device=torch.device("cuda", codec.device))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return codec.hist_read()
    return allreduce_counts(codec.hist_read().astype(np.int64), group).astype(np.uint64)[:q]


def distribution(counts):
    """prob = freq / sum(freq) (get_encoded_distribution.py:134), float64[q] as saved to distribution_info_N.npy."""
    c = np.asarray(counts, dtype=np.float64)
    return c / c.sum()
