#!/usr/bin/env python
"""Device-to-host copy bandwidth of every GPU of the box, alone and all together, with the pinned host buffers placed
(a) wherever the allocating thread happens to run (what bench.py's e2e did in round 1) and (b) on the NUMA node of the GPU
(ort_bind_host_thread before ort_host_alloc).  One process, one stream per GPU, 1 GiB copies through cudaMemcpyAsync
(torch), timed with CUDA events per GPU and wall clock for the aggregate.  Separates "the platform caps eight concurrent
device-to-host streams" from "our pinned allocations sit on the wrong socket".
usage: pcie_d2h_concurrent.py [GiB per copy = 1] [reps = 5]      -> one JSON object"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ort_b200 as ort  # noqa: E402

GIB = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nbytes = int(GIB * (1 << 30))
ng = torch.cuda.device_count()
ctxs = [ort.Context(d) for d in range(ng)]
cpus0 = sorted(os.sched_getaffinity(0))
out = {"gpus": ng, "copy_bytes": nbytes, "process_cpus": f"{cpus0[0]}-{cpus0[-1]} ({len(cpus0)})",
       "numa": [dict(zip(("node", "cpus"), c.device_numa())) for c in ctxs]}


def host_view(pa):
    return torch.from_numpy(pa.array)


def measure(bufs_h, bufs_d, streams, which):
    """copy on the GPUs in `which` at the same time; per-GPU GB/s from events, aggregate from the wall clock"""
    evs = {}
    for d in which:
        torch.cuda.set_device(d)
        with torch.cuda.stream(streams[d]):
            bufs_h[d].copy_(bufs_d[d], non_blocking=True)          # warm-up
    for d in which:
        streams[d].synchronize()
    t0 = time.perf_counter()
    for d in which:
        torch.cuda.set_device(d)
        with torch.cuda.stream(streams[d]):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(REPS):
                bufs_h[d].copy_(bufs_d[d], non_blocking=True)
            e1.record()
            evs[d] = (e0, e1)
    for d in which:
        streams[d].synchronize()
    wall = time.perf_counter() - t0
    per = {d: nbytes * REPS / (evs[d][0].elapsed_time(evs[d][1]) * 1e-3) / 1e9 for d in which}
    return per, nbytes * REPS * len(which) / wall / 1e9


streams, bufs_d = {}, {}
for d in range(ng):
    torch.cuda.set_device(d)
    streams[d] = torch.cuda.Stream(device=d)
    bufs_d[d] = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}")
for placement in ("default", "gpu_numa_node"):
    pinned, bound = [], []
    for d in range(ng):
        if placement == "gpu_numa_node":
            bound.append(ctxs[d].bind_host_thread())
        pa = ort.PinnedArray((nbytes,), dtype=np.uint8)
        pa.array[::4096] = 1                                        # first touch under the current policy
        pinned.append(pa)
    os.sched_setaffinity(0, cpus0)
    bufs_h = {d: host_view(pinned[d]) for d in range(ng)}
    alone = {}
    for d in range(ng):
        per, _ = measure(bufs_h, bufs_d, streams, [d])
        alone[d] = per[d]
    per_all, agg = measure(bufs_h, bufs_d, streams, list(range(ng)))
    half = list(range(ng // 2)) if ng > 1 else [0]
    _, agg_half = measure(bufs_h, bufs_d, streams, half)
    out[placement] = {"bound": bound, "alone_GBps": [round(alone[d], 1) for d in range(ng)],
                      "all_concurrent_per_gpu_GBps": [round(per_all[d], 1) for d in range(ng)],
                      "all_concurrent_aggregate_GBps": round(agg, 1), "first_half_concurrent_aggregate_GBps": round(agg_half, 1),
                      "sum_of_alone_GBps": round(sum(alone.values()), 1)}
    del bufs_h
    for pa in pinned:
        pa.free()
print(json.dumps(out))
