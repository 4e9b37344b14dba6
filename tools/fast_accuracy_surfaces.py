#!/usr/bin/env python
"""Per-surface POINTWISE error of FAST against STRICT arithmetic (STRICT is the mode the tests hold bit-identical to the CPU
oracle): the rays of the bench pupil grid (4.2 M per field, every 4th row and column of the 5792 x 2896 grid, 5 fields) go
through ort_trace3d_rays_dev with every surface recorded, and for each of the 12 loop steps the distribution of
    |x_fast - x_strict| / |x_strict|   and   |y_fast - y_strict| / |y_strict|       (pointwise, no position scale)
is histogrammed in decades -- beside the scale-relative figure the tests enforce (|delta| / max |coordinate| of the ray over
all surfaces).  Direction cosines: max |delta k|.  Prints one JSON object (-> profiles/r02_fast_vs_strict_per_surface.json)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ort_b200 as ort  # noqa: E402

NY, NX, SUB = 5792, 2896, 4
ctx = ort.Context(0)
ort.set_default_backend(ctx)
P = ort.prescriptions.DOUBLE_GAUSS
s = ort.solve(P["surfaces"], P["a"], P["h"])
Hs = ort.prescriptions.DOUBLE_GAUSS_FIELDS
p = ort.host._full_trace_setup(s.layout, s, Hs, 64, None, ctx)
ctx.set_layout(p["ext"], p["K"])
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
ns = p["ext"].shape[0] - 1
edges = [0.0, 1e-17, 1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-6, 1.0]
hist = np.zeros((ns, len(edges) - 1), dtype=np.int64)
worst_pt = np.zeros(ns); worst_sc = np.zeros(ns); small = np.zeros(ns, dtype=np.int64)
worst_k, nrays = 0.0, 0
xs = ort.host.jl_range(0.0, p["y_EP"], NX)[::SUB]
for j in range(len(Hs)):
    ys = ort.host.jl_range(p["y1"][j], p["y2"][j], NY)[::SUB]
    Y, X = np.meshgrid(ys, xs, indexing="ij")
    N = Y.size
    d = [torch.from_numpy(np.ascontiguousarray(a.reshape(-1))).to(dev) for a in (Y, X)]
    d += [torch.full((N,), float(p["u"][j]), dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)]
    res = {}
    for name, arith in (("fast", ort.FAST), ("strict", ort.STRICT)):
        xv = torch.empty((ns, N), dtype=torch.float64, device=dev); yv = torch.empty_like(xv)
        k = torch.empty((3, N), dtype=torch.float64, device=dev); fl = torch.empty(N, dtype=torch.uint8, device=dev)
        ctx.trace3d_rays_dev(N, *[t.data_ptr() for t in d], xv.data_ptr(), yv.data_ptr(), k.data_ptr(), fl.data_ptr(), arith=arith, stream=st)
        torch.cuda.synchronize()
        res[name] = (xv, yv, k, fl)
    (xf, yf, kf, ff), (xs_, ys_, ks, fs) = res["fast"], res["strict"]
    ok = (ff == 0) & (fs == 0)
    scale = torch.maximum(xs_.abs().amax(0), ys_.abs().amax(0)).clamp_min(1e-300)
    for c_f, c_s in ((xf, xs_), (yf, ys_)):
        dlt = (c_f - c_s).abs()
        for i in range(ns):
            ref = c_s[i].abs()
            sel = ok & (ref > 1e-6)                               # pointwise is meaningless for a coordinate that IS zero (x = 0 column)
            pt = (dlt[i][sel] / ref[sel])
            hist[i] += np.histogram(pt.cpu().numpy(), bins=edges)[0]
            worst_pt[i] = max(worst_pt[i], float(pt.max()) if pt.numel() else 0.0)
            worst_sc[i] = max(worst_sc[i], float((dlt[i][ok] / scale[ok]).max()))
            small[i] += int((ok & (ref <= 1e-6)).sum())
    worst_k = max(worst_k, float((kf - ks).abs()[:, ok].max()))
    nrays += int(ok.sum())
out = {"_what": __doc__.split("\n\n")[0] if False else "FAST vs STRICT per surface, pointwise and scale-relative; bench pupil grid subsampled 4 x 4, 5 fields",
       "rays_compared": nrays, "bin_edges": edges, "max_abs_direction_cosine_error": worst_k,
       "surfaces": [{"step": i + 1, "pointwise_max": worst_pt[i], "scale_relative_max": worst_sc[i],
                     "pointwise_decade_counts": hist[i].tolist(), "coordinates_below_1e-6_mm_excluded": int(small[i])} for i in range(ns)]}
print(json.dumps(out, indent=1))
