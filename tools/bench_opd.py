#!/usr/bin/env python
"""One 16 Mi-ray double-Gauss field with the OPD extension (ex, ey, opd, mask + OPD statistics; optionally the per-surface
apertures too), FAST arithmetic, timed with CUDA events: the A/B harness of the SIMPLE x EXT instantiation.
Prints one JSON object.  ORT_B200_LIB selects the library build."""
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
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
peak, _ = ctx.fp64_peak()
Pd = ort.prescriptions.DOUBLE_GAUSS
sd = ort.solve(Pd["surfaces"], Pd["a"], Pd["h"])
pe = ort.host._full_trace_setup(sd.layout, sd, [0.7], 64, None, ctx)
ys = torch.from_numpy(ort.host.jl_range(pe["y1"][0], pe["y2"][0], NY)).to(dev)
xs = torch.from_numpy(ort.host.jl_range(0.0, pe["y_EP"], NX)).to(dev)
st = torch.zeros(ort.STATS_BYTES, dtype=torch.uint8, device=dev)
bufs = {k: torch.empty(NY * NX, dtype=torch.float64, device=dev) for k in ("ex", "ey", "opd")}
bufs["mask"] = torch.empty(NY * NX, dtype=torch.uint8, device=dev)
fld = dict(u=float(pe["u"][0]), h_prime=float(pe["h_prime"][0]), opd_yc=float(pe["h_prime"][0]),
           opd_radius=float(pe["focus"] - sd.XP.t), opl_ref=150.0)
pt = {k: t.data_ptr() for k, t in bufs.items()}
pt["stats"] = st.data_ptr()
out = {"fp64_peak": peak}
Kc = np.where(np.isfinite(pe["ext"][:, 0]), -0.3, 0.0)
for name, K, ext, aps in (("opd_simple", pe["K"], ort.EXT_OPD, False), ("opd_vignette_simple", pe["K"], ort.EXT_OPD | ort.EXT_VIGNETTE, True),
                          ("opd_conics_general", Kc, ort.EXT_OPD, False)):
    ctx.set_layout(pe["ext"], K)
    ctx.set_apertures(np.append(Pd["a"], np.inf) if aps else None)

    def run():
        ctx.trace3d_grid_dev([fld], ys.data_ptr(), NY, xs.data_ptr(), NX, pe["stop"], pe["a_stop"], pt, stream=stream, ext=ext,
                             opd_scale=-1.0 / 587.5618e-6)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rec = np.frombuffer(st.cpu().numpy().tobytes(), dtype=ort.STATS_DTYPE)[0]
    out[name] = {"ms": ms, "fp64_frac_of_723": NY * NX * 723 / ms / 1e9 / peak, "kept": int(rec["n_kept"]), "n_vig": int(rec["n_vig"]),
                 "n_strict": int(rec["n_strict"]), "rms_opd_waves": float(np.sqrt(rec["m2_opd"] / max(int(rec["n_kept"]), 1)))}
print(json.dumps(out))
