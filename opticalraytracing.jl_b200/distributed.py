"""Multi-GPU sharding of the pupil-grid sweep: one process per GPU (torchrun), rays sharded by
contiguous blocks of y-rows (the outer loop index of src/PupilSampling.jl:123) so concatenating
the ranks' compacted outputs in rank order reproduces the reference's push! order, and ONE
all-gather of the per-field statistics records (80 B x fields per rank) -- the only exchange step
of the path.  Ray-level outputs stay on their GPU.  Backend: NCCL on GPUs, gloo in CPU tests.
"""
import numpy as np

from . import _lib
from .host import merge_stats, rms_from_stats


def shard_rows(ny_total, rank, world):
    """[lo, hi) y-rows of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(ny_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_stats(stats, group=None, device=None):
    """stats: structured array (n_fields,) of ort_stats on this rank -> (world, n_fields).
    Uses torch.distributed.all_gather_into_tensor on the raw bytes (NCCL needs a CUDA tensor: pass
    device='cuda')."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(stats).reshape(1, -1)
    world = dist.get_world_size(group)
    raw = np.ascontiguousarray(stats).view(np.uint8).reshape(-1)
    t = torch.from_numpy(raw.copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return np.frombuffer(out.cpu().numpy().tobytes(), dtype=_lib.STATS_DTYPE).reshape(world, -1)


def merged_field_stats(gathered):
    """(world, n_fields) records -> list of merged records, ranks folded in rank order (Chan), so every
    rank computes bit-identical results."""
    return [merge_stats(gathered[:, f]) for f in range(gathered.shape[1])]


def sharded_sweep(backend, fields, ys_total, xs, stop, a_stop, rank, world, group=None, device=None, **kw):
    """Trace this rank's block of y-rows and combine statistics across ranks.
    ys_total: (ny,) or (n_fields, ny).  Returns (local result dict, merged stats list, rms list)."""
    ys_total = np.asarray(ys_total, dtype=np.float64)
    lo, hi = shard_rows(ys_total.shape[-1], rank, world)
    res = backend.trace3d_grid(fields, ys_total[..., lo:hi], xs, stop, a_stop, **kw)
    gathered = allgather_stats(res["stats"], group=group, device=device)
    merged = merged_field_stats(gathered)
    return res, merged, [rms_from_stats(m) for m in merged]
