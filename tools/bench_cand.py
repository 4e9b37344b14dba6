#!/usr/bin/env python
"""K5 (BASELINE config 5) on one GPU: the per-candidate prelude (k_aim_candidates + k_aim_edges) and the aimed population
sweep (k_cand_classify + k_candidates_simple + k_candidates) timed separately with CUDA events; FAST arithmetic.
ORT_B200_LIB selects the library build.  usage: bench_cand.py [C]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ort_b200 as ort  # noqa: E402

C = int(float(sys.argv[1])) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
ctx = ort.Context(0)
st = torch.cuda.current_stream().cuda_stream
peak, _ = ctx.fp64_peak()
Pq = ort.prescriptions.COOKE
base = ort.prescriptions.perturbed_triplets(C)
d_R = torch.from_numpy(base).to(dev)
d_aim = torch.empty((C, 24), dtype=torch.float64, device=dev)
d_out = torch.empty((C, 4), dtype=torch.float64, device=dev)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"C": C, "fp64_peak": peak}
out["prelude_ms"] = timed(lambda: ctx.aim_candidates_dev(8, C, d_R.data_ptr(), Pq["a"], Pq["h"], 0.7, d_aim.data_ptr(), stream=st))
ms = timed(lambda: ctx.trace3d_candidates_aimed_dev(8, C, d_R.data_ptr(), d_aim.data_ptr(), 64, 64, d_out.data_ptr(), arith=ort.FAST, stream=st))
rays = C * 4096
tab = d_out.cpu().numpy()
out["sweep_aimed"] = {"ms": ms, "fp64_frac": rays * 513 / ms / 1e9 / peak, "rays_per_s": rays / ms * 1e3, "checksum": float(np.nansum(tab[:, 3]))}
# a mixed population: every 4th candidate gets a conic (general kernel)
mixed = base.copy()
mixed[::4, 3, 1] = -0.5
d_M = torch.from_numpy(mixed).to(dev)
ctx.aim_candidates_dev(8, C, d_M.data_ptr(), Pq["a"], Pq["h"], 0.7, d_aim.data_ptr(), aspheric=True, stream=st)
ms = timed(lambda: ctx.trace3d_candidates_aimed_dev(8, C, d_M.data_ptr(), d_aim.data_ptr(), 64, 64, d_out.data_ptr(), arith=ort.FAST, stream=st))
out["sweep_aimed_quarter_conic"] = {"ms": ms, "fp64_frac": rays * 513 / ms / 1e9 / peak}
# the bit-identical arithmetic on the first 8192 candidates of the population (one CTA per candidate, STRICT)
Cs = min(C, 8192)
ctx.aim_candidates_dev(8, Cs, d_R.data_ptr(), Pq["a"], Pq["h"], 0.7, d_aim.data_ptr(), stream=st)
ms = timed(lambda: ctx.trace3d_candidates_aimed_dev(8, Cs, d_R.data_ptr(), d_aim.data_ptr(), 64, 64, d_out.data_ptr(), arith=ort.STRICT, stream=st), reps=5, warm=2)
out["sweep_aimed_strict_8192"] = {"ms": ms, "fp64_frac": Cs * 4096 * 513 / ms / 1e9 / peak}
print(json.dumps(out))
