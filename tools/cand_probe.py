"""Probe: unaimed vs aimed candidate sweeps on the same population (for ncu metric comparison)."""
import sys
import numpy as np
import torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ort_b200 as ort

C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = ort.Context()
ort.set_default_backend(ctx)
dev = torch.device("cuda:0")
P = ort.prescriptions.COOKE
sysm = ort.solve(P["surfaces"], P["a"], P["h"])
p = ort.host._full_trace_setup(sysm.layout, sysm, [0.7], 64, None, ctx)
base = ort.prescriptions.perturbed_triplets(C)
rows = 9
RtnK = np.zeros((C, 4, rows)); RtnK[:, :, :-1] = base
RtnK[:, 0, -1] = np.inf; RtnK[:, 2, -1] = 1.0; RtnK[:, 1, -2] = p["focus"]
d_R = torch.from_numpy(RtnK).to(dev)
ys = torch.from_numpy(np.linspace(p["y1"][0], p["y2"][0], 64)).to(dev)
xs = torch.from_numpy(np.linspace(0.0, p["y_EP"], 64)).to(dev)
d_o = torch.empty((C, 4), dtype=torch.float64, device=dev)
fld = dict(u=float(p["u"][0]), v=0.0, h_prime=float(p["h_prime"][0]))
d_Rn = torch.from_numpy(base.copy()).to(dev)
d_aim = torch.empty((C, 24), dtype=torch.float64, device=dev)
d_oa = torch.empty((C, 4), dtype=torch.float64, device=dev)
ctx.aim_candidates_dev(8, C, d_Rn.data_ptr(), P["a"], P["h"], 0.7, d_aim.data_ptr())
for arith in (ort.FAST, ort.STRICT):
    ctx.trace3d_candidates_dev(rows, C, d_R.data_ptr(), fld, ys.data_ptr(), 64, xs.data_ptr(), 64, p["stop"], p["a_stop"], d_o.data_ptr(), arith=arith)
    ctx.trace3d_candidates_aimed_dev(8, C, d_Rn.data_ptr(), d_aim.data_ptr(), 64, 64, d_oa.data_ptr(), arith=arith)
    ctx.sync()
    a, b = d_o.cpu().numpy(), d_oa.cpu().numpy()
    print("arith", arith, "kept unaimed", a[:, 0].mean(), "aimed", b[:, 0].mean(), "rms", np.nanmean(a[:, 3]), np.nanmean(b[:, 3]))
