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


def sharded_candidates(backend, RtnK, a, h_prime, H, k_rays=64, rank=0, world=1, group=None, device=None, arith=_lib.FAST,
                       gather=True):
    """BASELINE config 5 across ranks: candidates are independent, so every rank takes a contiguous range
    [lo, hi) of the population (prescriptions replicated by the caller or sliced before the call) and runs the
    per-candidate prelude + aimed sweep on it -- no data-path collective.  With gather=True the (C, 4) merit table
    (n_kept, mean_x, mean_y, RMS; 32 B per candidate) is all-gathered so every rank can rank the whole population.
    Returns (table, (lo, hi)): the full table when gathered, else this rank's rows."""
    import torch
    import torch.distributed as dist
    RtnK = np.asarray(RtnK, dtype=np.float64)
    C = RtnK.shape[0]
    lo, hi = shard_rows(C, rank, world)
    aim = backend.aim_candidates(RtnK[lo:hi], a, h_prime, H)
    spot = backend.trace3d_candidates_aimed(RtnK[lo:hi], aim, int(k_rays), int(k_rays) // 2, arith=arith)
    if not gather or world == 1 or not (dist.is_available() and dist.is_initialized()):
        return spot, (lo, hi)
    per = -(-C // world)                                     # ranges differ by at most one row: pad to the longest
    mine = torch.full((per, 4), float("nan"), dtype=torch.float64)
    mine[:hi - lo] = torch.from_numpy(np.ascontiguousarray(spot))
    if device is not None:
        mine = mine.to(device)
    out = torch.empty((world, per, 4), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=group)
    out = out.cpu().numpy()
    table = np.concatenate([out[r, :shard_rows(C, r, world)[1] - shard_rows(C, r, world)[0]] for r in range(world)])
    return table, (lo, hi)
