import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ort_b200 as ort
ctx = ort.Context(); ort.set_default_backend(ctx)
P = ort.prescriptions.COOKE
sysm = ort.solve(P["surfaces"], P["a"], P["h"])
p = ort.host._full_trace_setup(sysm.layout, sysm, [0.7], 64, None, ctx)
base = ort.prescriptions.perturbed_triplets(16)
aim = ctx.aim_candidates(base, P["a"], P["h"], 0.7)
for c in range(6):
    y1, y2, y_EP, u, hp, focus, stop, a_stop = aim[c, :8]
    ext = np.vstack([base[c, :3].T, [np.inf, 0.0, 1.0]]); ext[-2, 1] = focus
    ctx.set_layout(ext)
    r = ctx.trace3d_grid([dict(u=u, v=0.0, h_prime=hp)], np.linspace(y1, y2, 64), np.linspace(0, y_EP, 64), int(stop), a_stop, arith=ort.FAST, want=("stats","flags"))
    s = r["stats"][0]
    ext[-2, 1] = p["focus"]; ctx.set_layout(ext)
    r2 = ctx.trace3d_grid([dict(u=float(p["u"][0]), v=0.0, h_prime=float(p["h_prime"][0]))], np.linspace(p["y1"][0], p["y2"][0], 64), np.linspace(0, p["y_EP"], 64), p["stop"], p["a_stop"], arith=ort.FAST, want=("stats","flags"))
    s2 = r2["stats"][0]
    print(c, "aimed: kept", s["n_kept"], "strict", s["n_strict"], "miss", s["n_miss"], "tir", s["n_tir"], "clip", s["n_clip"], "| nominal: kept", s2["n_kept"], "strict", s2["n_strict"], "miss", s2["n_miss"], "tir", s2["n_tir"])
    fl = r["flags"][0].reshape(64, 64)
    print("   aimed y-range", y1, y2, "nominal", p["y1"][0], p["y2"][0], " y_EP", y_EP, p["y_EP"])
