#!/usr/bin/env python
"""One 16 Mi-ray double-Gauss field (spot + mask, FAST arithmetic) timed for three variants of the same prescription,
so that every surface body of fast_step is measured, not just the one the bench workload uses:
  spheres   nominal (refracting spheres take the division-free body)
  conics    every curved surface given K = -0.3 (conic body)
  weak      nominal lens plus a zero-power, |R| = 1e5 mm dummy refracting pair in front (spheres with the division)
  poly      conics plus y^4, y^6, y^8 terms on every curved surface (EXTENSION: fast_step's polynomial body); timed in STRICT too
Prints one JSON object {variant: ms}.  ORT_B200_LIB selects the library build."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ort_b200 as ort  # noqa: E402

NY, NX = 5792, 2896
ctx = ort.Context(0)
ort.set_default_backend(ctx)
P = ort.prescriptions.DOUBLE_GAUSS
s = ort.solve(P["surfaces"], P["a"], P["h"])
p = ort.host._full_trace_setup(s.layout, s, (0.7,), 64, None, ctx)
dev = torch.device("cuda", 0)
xs = torch.from_numpy(ort.host.jl_range(0.0, p["y_EP"], NX)).to(dev)
ys = torch.from_numpy(ort.host.jl_range(p["y1"][0], p["y2"][0], NY)).to(dev)
b = {k: torch.empty(NY * NX, dtype=torch.float64, device=dev) for k in ("ex", "ey")}
b["mask"] = torch.empty(NY * NX, dtype=torch.uint8, device=dev)
st = torch.zeros(ort.STATS_BYTES, dtype=torch.uint8, device=dev)
ptrs = {k: v.data_ptr() for k, v in b.items()}
ptrs["stats"] = st.data_ptr()
ext = np.array(p["ext"], dtype=np.float64)
K0 = np.zeros(len(ext))
variants = {"spheres": (ext, K0, p["stop"], None, ort.FAST)}
Kc = np.where(np.isfinite(ext[:, 0]), -0.3, 0.0)
variants["conics"] = (ext, Kc, p["stop"], None, ort.FAST)
weak = np.vstack([ext[:1], [[1e5, 1.0, 1.5], [1e5, 1.0, 1.0]], ext[1:]])
variants["weak"] = (weak, np.zeros(len(weak)), p["stop"] + 2, None, ort.FAST)
Pc = np.zeros((len(ext), 9))
Pc[np.isfinite(ext[:, 0]), 4], Pc[np.isfinite(ext[:, 0]), 6], Pc[np.isfinite(ext[:, 0]), 8] = 2e-7, -1e-10, 5e-14
Pc[0] = 0.0
variants["poly"] = (ext, Kc, p["stop"], Pc, ort.FAST)
variants["poly_strict"] = (ext, Kc, p["stop"], Pc, ort.STRICT)
out = {}
stream = torch.cuda.current_stream().cuda_stream
for name, (M, K, stop, poly, arith) in variants.items():
    ctx.set_layout(M, K)
    ctx.set_polynomials(poly)
    fld = [dict(u=float(p["u"][0]), h_prime=float(p["h_prime"][0]))]
    for _ in range(3):
        ctx.trace3d_grid_dev(fld, ys.data_ptr(), NY, xs.data_ptr(), NX, stop, p["a_stop"], ptrs, stream=stream, arith=arith)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ctx.trace3d_grid_dev(fld, ys.data_ptr(), NY, xs.data_ptr(), NX, stop, p["a_stop"], ptrs, stream=stream, arith=arith)
    e1.record()
    torch.cuda.synchronize()
    rec = np.frombuffer(st.cpu().numpy().tobytes(), dtype=ort.STATS_DTYPE)[0]
    out[name] = {"ms": e0.elapsed_time(e1) / 10, "kept": int(rec["n_kept"]), "n_strict": int(rec["n_strict"])}
    if name == "poly":
        keep = {k: v.clone() for k, v in b.items()}
    if name == "poly_strict":                     # FAST against the reference arithmetic on all 16 Mi rays
        m = b["mask"] != 0
        d = torch.maximum((keep["ex"] - b["ex"]).abs()[m].max(), (keep["ey"] - b["ey"]).abs()[m].max())
        out[name]["fast_vs_strict"] = {"mask_xor": int((keep["mask"] != b["mask"]).sum()), "max_abs_mm": float(d),
                                       "max_rel_to_position_scale": float(d) / max(abs(float(p["h_prime"][0])), 1.0)}
print(json.dumps(out))
