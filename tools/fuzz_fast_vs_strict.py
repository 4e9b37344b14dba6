"""Fuzz the FAST path against STRICT on the GPU: seeded random sequential systems (spheres, planes, conics, mirrors,
cemented groups; tests/test_gpu_random_systems.random_system) and deliberately hostile ray bundles (large heights ->
misses and near-equator hits, steep slopes -> TIR and grazing incidence).  STRICT is bit-identical to the CPU restatement
of the reference (tests), so agreement here is agreement with the reference: flags and NaN patterns must be IDENTICAL,
positions within 1e-12 of the position scale on well-conditioned rays (outliers are re-examined against the 80-bit oracle).
usage: python tools/fuzz_fast_vs_strict.py [n_systems] [rays_per_system] [poly]  -> JSON summary on stdout
       python tools/fuzz_fast_vs_strict.py grid [n_systems] [poly]
`poly`: refracting systems only, with random y^3 .. y^8 terms on most curved surfaces (EXTENSION, fast_step's polynomial body);
the 80-bit oracle has no polynomial terms, so outliers are counted but not re-examined."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import ort_b200 as ort
from test_gpu_random_systems import random_system


def random_terms(rng, S):
    """coefficient rows for random_system's S: y^4, y^6, y^8 (and sometimes y^3) terms on ~70 % of the curved surfaces"""
    rows = S.shape[0]
    P = np.zeros((rows, 9))
    for i in range(1, rows):
        if np.isfinite(S[i, 0]) and rng.uniform() < 0.7:
            P[i, 4], P[i, 6], P[i, 8] = rng.uniform(-3e-7, 3e-7), rng.uniform(-3e-10, 3e-10), rng.uniform(-1e-13, 1e-13)
            if rng.uniform() < 0.3: P[i, 3] = rng.uniform(-1e-6, 1e-6)
    return P


def grid_mode(nsys, poly=False):
    """FAST vs STRICT through the grid kernel: random systems, random stop surface and radius, collimated and finite-object
    fields, every output-set instantiation (generic / spot + mask / statistics only): masks, flags and counts identical."""
    ctx = ort.Context(0)
    tot = dict(mode="grid" + ("+poly" if poly else ""), systems=nsys, rays=0, mask_mismatch=0, flag_mismatch=0, count_mismatch=0, lean_mismatch=0,
               stats_only_mismatch=0, max_err=0.0, kept=0, strict_retraced=0)
    for seed in range(nsys):
        rng = np.random.default_rng(70000 + seed)
        S = random_system(rng, mirrors=seed % 3 == 1 and not poly, conics=seed % 2 == 1)
        K = np.append(S[:, 3], 0.0)
        ext = np.vstack([S[:, :3], [np.inf, 0.0, S[-1, 2]]])
        ext[-2, 1] = rng.uniform(20.0, 80.0) * np.sign(S[-1, 2])
        stop = int(rng.integers(1, ext.shape[0] - 1))
        a_stop = rng.uniform(3.0, 12.0)
        ny, nx = int(rng.integers(40, 90)), int(rng.integers(20, 50))
        ys, xs = np.linspace(-14, 14, ny), np.linspace(0, 14, nx)
        if seed % 4 == 3:
            fld = dict(mode=1, ybar=float(rng.uniform(-20, 20)), z0=float(-rng.uniform(150, 600)), h_prime=0.3)
        else:
            fld = dict(u=float(rng.uniform(-0.2, 0.2)), v=float(rng.uniform(-0.05, 0.05)), h_prime=0.3)
        ctx.set_layout(ext, K)
        if poly: ctx.set_polynomials(np.vstack([random_terms(rng, S), np.zeros((1, 9))]))
        want = ("ex", "ey", "r", "theta", "mask", "flags", "stats")
        rs = ctx.trace3d_grid([fld], ys, xs, stop, a_stop, arith=ort.STRICT, want=want)
        rf = ctx.trace3d_grid([fld], ys, xs, stop, a_stop, arith=ort.FAST, want=want)
        rl = ctx.trace3d_grid([fld], ys, xs, stop, a_stop, arith=ort.FAST, want=("ex", "ey", "mask", "stats"))
        ro = ctx.trace3d_grid([fld], ys, xs, stop, a_stop, arith=ort.FAST, want=("stats",))
        tot["rays"] += ny * nx
        tot["mask_mismatch"] += int(np.count_nonzero(rs["mask"] != rf["mask"]))
        tot["flag_mismatch"] += int(np.count_nonzero(rs["flags"] != rf["flags"]))
        tot["count_mismatch"] += int(rs["stats"]["n_kept"][0] != rf["stats"]["n_kept"][0])
        tot["lean_mismatch"] += int(np.count_nonzero(rl["mask"] != rf["mask"])) + int(rl["stats"].tobytes() != rf["stats"].tobytes())
        for k in rl["stats"].dtype.names:           # which statistics differ between the output-set instantiations, if any
            if rl["stats"][k].tobytes() != rf["stats"][k].tobytes():
                tot.setdefault("lean_stat_fields", {}).setdefault(k, 0)
                tot["lean_stat_fields"][k] += 1
        tot["lean_ex_bits"] = tot.get("lean_ex_bits", 0) + int(np.count_nonzero(rl["ex"][0].view(np.int64) != rf["ex"][0].view(np.int64)))
        tot["stats_only_mismatch"] += int(ro["stats"].tobytes() != rf["stats"].tobytes())
        tot["kept"] += int(rs["stats"]["n_kept"][0]); tot["strict_retraced"] += int(rf["stats"]["n_strict"][0])
        m = rs["mask"][0].astype(bool)
        if m.any():
            sc = max(np.abs(rs["ex"][0][m]).max(), np.abs(rs["ey"][0][m]).max(), 10.0)
            e = np.maximum(np.abs(rf["ex"][0] - rs["ex"][0]), np.abs(rf["ey"][0] - rs["ey"][0])) / sc
            e[~m] = 0.0
            j = int(np.argmax(e))
            if float(e[j]) > tot["max_err"]:
                tot["max_err"] = float(e[j])
                tot["worst"] = dict(seed=seed, iy=j // nx, ix=j % nx, y0=float(ys[j // nx]), x0=float(xs[j % nx]), err=float(e[j]),
                                    field=fld, stop=stop, a_stop=float(a_stop), r_over_a=float(rs["r"][0][j] / a_stop))
            tot["over_1e12"] = tot.get("over_1e12", 0) + int(np.count_nonzero(e > 1e-12))
    print(json.dumps(tot, indent=1))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "grid":
        return grid_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 300, poly="poly" in sys.argv[3:])
    poly = "poly" in sys.argv[3:]
    nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    ctx = ort.Context(0)
    try:
        from oracle import oracle as orc
    except Exception:
        orc = None
    if poly: orc = None
    tot = dict(mode="rays" + ("+poly" if poly else ""), systems=nsys, rays=0, flag_mismatch=0, nan_mismatch=0, miss=0, tir=0, domain=0, finite=0, over_1e12=0,
               over_1e12_well_conditioned=0, max_err=0.0, max_err_well_conditioned=0.0)
    worst = []
    for seed in range(nsys):
        rng = np.random.default_rng(50000 + seed)
        S = random_system(rng, mirrors=seed % 3 == 1 and not poly, conics=seed % 2 == 1)
        K = S[:, 3].copy()
        hostile = seed % 4
        h = (9.0, 25.0, 40.0, 15.0)[hostile]
        sl = (0.12, 0.3, 0.15, 0.6)[hostile]
        y0, x0 = rng.uniform(-h, h, N), rng.uniform(-h, h, N)
        u0, v0 = rng.uniform(-sl, sl, N), rng.uniform(-sl, sl, N)
        ctx.set_layout(S[:, :3], K)
        if poly: ctx.set_polynomials(random_terms(rng, S))
        xs, ys, ks, fs = ctx.trace3d_rays(y0, x0, u0, v0, arith=ort.STRICT)
        xf, yf, kf, ff = ctx.trace3d_rays(y0, x0, u0, v0, arith=ort.FAST)
        tot["rays"] += N
        tot["flag_mismatch"] += int(np.count_nonzero(fs != ff))
        tot["nan_mismatch"] += int(np.count_nonzero(np.isnan(xs) != np.isnan(xf)) + np.count_nonzero(np.isnan(ys) != np.isnan(yf)))
        tot["miss"] += int(np.count_nonzero(fs & ort.FLAG_MISS)); tot["tir"] += int(np.count_nonzero(fs & ort.FLAG_TIR))
        tot["domain"] += int(np.count_nonzero(fs & ort.FLAG_DOMAIN))
        with np.errstate(all="ignore"):
            scale = np.maximum(np.nan_to_num(np.nanmax(np.abs(np.stack([xs, ys])), axis=(0, 1)), nan=1.0), 1.0)
            e = np.maximum(np.nan_to_num(np.nanmax(np.abs(xf - xs), axis=0), nan=0.0),
                           np.nan_to_num(np.nanmax(np.abs(yf - ys), axis=0), nan=0.0)) / scale
        tot["finite"] += int(np.count_nonzero(~np.isnan(xs[-1])))
        bad = np.nonzero(e > 1e-12)[0]
        tot["over_1e12"] += len(bad)
        tot["max_err"] = max(tot["max_err"], float(e.max()))
        if len(bad) and orc is not None:          # is STRICT itself that far from the 80-bit truth on these rays?
            xl, yl, _ = orc.trace3d_ld_batch(S[:, :3], y0[bad], x0[bad], u0[bad], v0[bad], K=K)
            with np.errstate(all="ignore"):
                cond = np.maximum(np.nan_to_num(np.nanmax(np.abs(xs[:, bad] - xl), axis=0), nan=0.0),
                                  np.nan_to_num(np.nanmax(np.abs(ys[:, bad] - yl), axis=0), nan=0.0)) / scale[bad]
                etrue = np.maximum(np.nan_to_num(np.nanmax(np.abs(xf[:, bad] - xl), axis=0), nan=0.0),
                                   np.nan_to_num(np.nanmax(np.abs(yf[:, bad] - yl), axis=0), nan=0.0)) / scale[bad]
            well = cond < 1e-13
            tot["over_1e12_well_conditioned"] += int(np.count_nonzero(well))
            for j in np.nonzero(etrue > 5e-13 + 2 * cond)[0]:
                worst.append(dict(seed=seed, ray=int(bad[j]), err_vs_strict=float(e[bad[j]]), strict_vs_truth=float(cond[j]),
                                  fast_vs_truth=float(etrue[j])))
        ewell = e.copy(); ewell[bad] = 0.0
        tot["max_err_well_conditioned"] = max(tot["max_err_well_conditioned"], float(ewell.max()))
    tot["fast_worse_than_strict_vs_truth"] = worst[:20]
    tot["n_fast_worse_than_strict_vs_truth"] = len(worst)
    print(json.dumps(tot, indent=1))


if __name__ == "__main__":
    main()
