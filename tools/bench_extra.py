#!/usr/bin/env python
"""Measures the secondary kernels of the hot path on one B200 (not a driver bench line):
  K3 paraxial y-nu trace, 40-row Lens (BASELINE config 4 shape)     -- HBM / FP64 balanced
  K4 transfer-matrix apply (forward and reverse)                    -- HBM bound
  K5 candidate-batched 3-D trace (BASELINE config 5 shape)          -- FP64 bound
  K1 with all outputs (r, theta) and with ordered compaction
Prints one JSON object; device-resident inputs, CUDA-event timing via ort_profile_*."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ort_b200 as ort  # noqa: E402

HBM = 6544.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timed(ctx, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    t = ctx.profile_read()
    ctx.profile_enable(False)
    return float(np.mean(t)), float(np.min(t))


def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28          # rays for K3 / K4
    dev = torch.device("cuda", 0)
    ctx = ort.Context(0)
    ort.set_default_backend(ctx)
    st = torch.cuda.current_stream().cuda_stream
    peak, _ = ctx.fp64_peak()
    out = {"fp64_peak_tflops": peak, "hbm_peak_gbs": HBM, "N": N}

    # ---- K3: 40-row paraxial trace, rays generated on the device (counter-based, seed 42)
    S = ort.prescriptions.zoom20()
    lens = ort.make_lens(S)
    g = torch.Generator(device=dev); g.manual_seed(42)
    y0 = (torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 20 - 10)
    w0 = (torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 0.4 - 0.2)
    y = torch.empty_like(y0); w = torch.empty_like(w0)
    ci = torch.empty(N, dtype=torch.int32, device=dev)
    a = np.full(len(lens.tau), 25.0)
    for name, arith, clip in (("fast", ort.FAST, False), ("strict", ort.STRICT, False), ("fast_clip", ort.FAST, True)):
        ms, best = timed(ctx, lambda: ctx.paraxial_batch_dev(lens.tau, lens.phi, N, y0.data_ptr(), w0.data_ptr(), y.data_ptr(),
                                                            w.data_ptr(), ci.data_ptr() if clip else None, a=a if clip else None,
                                                            clip=clip, arith=arith, stream=st))
        byts = N * (32 + (4 if clip else 0))
        out[f"paraxial40_{name}"] = {"ms": ms, "rays_per_s": N / ms * 1e3, "GBps": byts / ms / 1e6, "hbm_frac": byts / ms / 1e6 / HBM,
                                     "tflops_nominal": N * 160 / ms / 1e9, "fp64_frac": N * 160 / ms / 1e9 / peak}
    # ---- K4: transfer matrix
    sysm = ort.solve(ort.prescriptions.COOKE["surfaces"], ort.prescriptions.COOKE["a"], ort.prescriptions.COOKE["h"])
    v = torch.stack([y0, w0], dim=1).contiguous()
    vo = torch.empty_like(v)
    for name, rev in (("forward", False), ("reverse", True)):
        ms, best = timed(ctx, lambda: ctx.transfer_batch_dev(sysm.M, -50.0, 77.4, N, v.data_ptr(), vo.data_ptr(), reverse=rev, stream=st))
        out[f"transfer_{name}"] = {"ms": ms, "rays_per_s": N / ms * 1e3, "GBps": N * 32 / ms / 1e6, "hbm_frac": N * 32 / ms / 1e6 / HBM}
    # size-independent property at full size: reverse(forward(v)) == v
    ctx.transfer_batch_dev(sysm.M, -50.0, 77.4, N, v.data_ptr(), vo.data_ptr(), stream=st)
    back = torch.empty_like(v)
    ctx.transfer_batch_dev(sysm.M, -50.0, 77.4, N, vo.data_ptr(), back.data_ptr(), reverse=True, stream=st)
    torch.cuda.synchronize()
    out["transfer_roundtrip_max_abs_err"] = float((back - v).abs().max())
    del v, vo, back, y0, w0, y, w, ci
    torch.cuda.empty_cache()

    # ---- k_rays: arbitrary rays, every surface recorded (raytrace(surfaces, y, x, U, V, Vector{RealRay})), device-resident
    Pg = ort.prescriptions.DOUBLE_GAUSS
    ctx.set_layout(Pg["surfaces"])
    NR = 1 << 24
    gr = torch.Generator(device=dev); gr.manual_seed(3)
    ry = torch.rand(NR, dtype=torch.float64, device=dev, generator=gr) * 20 - 10
    rx = torch.rand(NR, dtype=torch.float64, device=dev, generator=gr) * 20 - 10
    ru = torch.rand(NR, dtype=torch.float64, device=dev, generator=gr) * 0.2 - 0.1
    rv = torch.rand(NR, dtype=torch.float64, device=dev, generator=gr) * 0.2 - 0.1
    ns = Pg["surfaces"].shape[0] - 1
    oxv = torch.empty((ns, NR), dtype=torch.float64, device=dev); oyv = torch.empty_like(oxv)
    ok_ = torch.empty((3, NR), dtype=torch.float64, device=dev); ofl = torch.empty(NR, dtype=torch.uint8, device=dev)
    for name, arith in (("fast", ort.FAST), ("strict", ort.STRICT)):
        ms, best = timed(ctx, lambda: ctx.trace3d_rays_dev(NR, ry.data_ptr(), rx.data_ptr(), ru.data_ptr(), rv.data_ptr(), oxv.data_ptr(),
                                                          oyv.data_ptr(), ok_.data_ptr(), ofl.data_ptr(), arith=arith, stream=st), reps=5, warm=2)
        bytes_ray = 32 + ns * 16 + 24 + 1
        out[f"rays_all_surfaces_{name}"] = {"rays": NR, "ms": ms, "rays_per_s": NR / ms * 1e3, "bytes_per_ray": bytes_ray,
                                            "GBps": NR * bytes_ray / ms / 1e6, "hbm_frac": NR * bytes_ray / ms / 1e6 / HBM}
    del ry, rx, ru, rv, oxv, oyv, ok_, ofl
    torch.cuda.empty_cache()
    # ---- K5: candidates (config 5: C triplet variants x 4096 rays)
    C = int(float(sys.argv[2])) if len(sys.argv) > 2 else 65536
    p = ort.host._full_trace_setup(sysm.layout, sysm, [0.7], 64, None, ctx)
    rows = p["ext"].shape[0]
    base = ort.prescriptions.perturbed_triplets(C)
    RtnK = np.zeros((C, 4, rows)); RtnK[:, :, :-1] = base
    RtnK[:, 0, -1] = np.inf; RtnK[:, 2, -1] = 1.0; RtnK[:, 1, -2] = p["focus"]
    d_R = torch.from_numpy(RtnK).to(dev)
    ys = torch.from_numpy(np.linspace(p["y1"][0], p["y2"][0], 64)).to(dev)
    xs = torch.from_numpy(np.linspace(-p["y_EP"], p["y_EP"], 64)).to(dev)
    d_o = torch.empty((C, 4), dtype=torch.float64, device=dev)
    fld = dict(u=float(p["u"][0]), v=0.0, h_prime=float(p["h_prime"][0]))
    for name, arith in (("fast", ort.FAST), ("strict", ort.STRICT)):
        ms, best = timed(ctx, lambda: ctx.trace3d_candidates_dev(rows, C, d_R.data_ptr(), fld, ys.data_ptr(), 64, xs.data_ptr(), 64,
                                                                p["stop"], p["a_stop"], d_o.data_ptr(), arith=arith, stream=st), reps=5, warm=2)
        rays = C * 4096
        out[f"candidates_{name}"] = {"ms": ms, "candidates_per_s": C / ms * 1e3, "rays_per_s": rays / ms * 1e3,
                                     "intersections_per_s_x7": rays * 7 / ms * 1e3, "tflops_nominal": rays * 513 / ms / 1e9,
                                     "fp64_frac": rays * 513 / ms / 1e9 / peak}
    # ---- f1: per-candidate prelude (solve + aimed chief / marginal / edge rays) and the aimed population sweep
    d_Rn = torch.from_numpy(base.copy()).to(dev)
    d_aim = torch.empty((C, 24), dtype=torch.float64, device=dev)
    d_oa = torch.empty((C, 4), dtype=torch.float64, device=dev)
    Pq = ort.prescriptions.COOKE
    ms, best = timed(ctx, lambda: ctx.aim_candidates_dev(8, C, d_Rn.data_ptr(), Pq["a"], Pq["h"], 0.7, d_aim.data_ptr(), stream=st), reps=5, warm=2)
    out["aim_candidates"] = {"ms": ms, "candidates_per_s": C / ms * 1e3, "failed": int((d_aim[:, 11] != 0).sum())}
    for name, arith in (("fast", ort.FAST), ("strict", ort.STRICT)):
        ms, best = timed(ctx, lambda: ctx.trace3d_candidates_aimed_dev(8, C, d_Rn.data_ptr(), d_aim.data_ptr(), 64, 64, d_oa.data_ptr(),
                                                                      arith=arith, stream=st), reps=5, warm=2)
        rays = C * 4096
        out[f"candidates_aimed_{name}"] = {"ms": ms, "candidates_per_s": C / ms * 1e3, "rays_per_s": rays / ms * 1e3,
                                           "fp64_frac": rays * 513 / ms / 1e9 / peak}
    ra = d_oa.cpu().numpy()
    out["candidates_aimed_rms_range"] = [float(np.nanmin(ra[:, 3])), float(np.nanmax(ra[:, 3]))]
    # ---- K7: first-order solve + Seidel sums per candidate
    d_s = torch.empty((C, 16), dtype=torch.float64, device=dev)
    Pc = ort.prescriptions.COOKE
    d_Rs = torch.from_numpy(ort.prescriptions.perturbed_triplets(C)).to(dev)
    ms, best = timed(ctx, lambda: ctx.seidel_candidates_dev(8, C, d_Rs.data_ptr(), Pc["a"], Pc["h"], d_s.data_ptr(), stream=st), reps=5, warm=2)
    out["seidel_candidates"] = {"ms": ms, "candidates_per_s": C / ms * 1e3, "GBps": C * (8 * 4 * 8 + 128) / ms / 1e6}

    # ---- config 1 (the reference's own CPU-runnable case): Cooke triplet, 64 x 64 pupil grid, 3 fields, spot RMS,
    #      end to end through the public API (host prelude + ray aiming + sweep + compaction + D2H + mirror)
    import time
    ort.full_trace_fields(sysm.layout, sysm, [0.0, 0.7, 1.0], 64)
    l0 = ctx.launch_count(); t0 = time.perf_counter()
    for _ in range(20):
        errs = ort.full_trace_fields(sysm.layout, sysm, [0.0, 0.7, 1.0], 64)
    dt = (time.perf_counter() - t0) / 20
    out["config1_full_trace_3_fields"] = {"ms": dt * 1e3, "launches_per_call": (ctx.launch_count() - l0) / 20,
                                          "rays": 3 * 2048, "rms": [e.RMS for e in errs]}
    try:
        from oracle import prelude as pre
        so = pre.solve(Pc["surfaces"], Pc["a"], Pc["h"])
        t0 = time.perf_counter()
        for _ in range(3):
            ref = [pre.full_trace(so, H, 64) for H in (0.0, 0.7, 1.0)]
        out["config1_cpu_oracle_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    except Exception as e:      # the oracle is optional here
        out["config1_cpu_oracle_ms"] = str(e)

    # ---- extension kernels: OPD accumulation (+ per-surface apertures) on the bench grid, one field
    torch.cuda.empty_cache()
    Pd = ort.prescriptions.DOUBLE_GAUSS
    sd = ort.solve(Pd["surfaces"], Pd["a"], Pd["h"])
    pe = ort.host._full_trace_setup(sd.layout, sd, [0.7], 64, None, ctx)
    ctx.set_layout(pe["ext"], pe["K"])
    ctx.set_apertures(np.append(Pd["a"], np.inf))
    nyb, nxb = 5792, 2896
    ysb = torch.from_numpy(np.linspace(pe["y1"][0], pe["y2"][0], nyb)).to(dev)
    xsb = torch.from_numpy(np.linspace(0.0, pe["y_EP"], nxb)).to(dev)
    stb = torch.zeros(ort.STATS_BYTES, dtype=torch.uint8, device=dev)
    bufs = {k: torch.empty(nyb * nxb, dtype=torch.float64, device=dev) for k in ("ex", "ey", "opd")}
    bufs["mask"] = torch.empty(nyb * nxb, dtype=torch.uint8, device=dev)
    fldb = dict(u=float(pe["u"][0]), h_prime=float(pe["h_prime"][0]), opd_yc=float(pe["h_prime"][0]), opd_radius=float(pe["focus"] - sd.XP.t),
                opl_ref=150.0)
    pt = {k: v.data_ptr() for k, v in bufs.items()}; pt["stats"] = stb.data_ptr()
    for name, ext in (("opd", ort.EXT_OPD), ("opd_vignette", ort.EXT_OPD | ort.EXT_VIGNETTE)):
        ms, best = timed(ctx, lambda: ctx.trace3d_grid_dev([fldb], ysb.data_ptr(), nyb, xsb.data_ptr(), nxb, pe["stop"], pe["a_stop"], pt,
                                                          stream=st, ext=ext, opd_scale=-1.0 / 587.5618e-6), reps=5, warm=2)
        rec = np.frombuffer(stb.cpu().numpy().tobytes(), dtype=ort.STATS_DTYPE)[0]
        out[f"grid_ext_{name}"] = {"rays": nyb * nxb, "ms": ms, "rays_per_s": nyb * nxb / ms * 1e3, "fp64_frac_of_723": nyb * nxb * 723 / ms / 1e9 / peak,
                                   "kept": int(rec["n_kept"]), "n_vig": int(rec["n_vig"]), "rms_opd_waves": float(np.sqrt(rec["m2_opd"] / max(rec["n_kept"], 1)))}
    ctx.set_apertures(None)
    del bufs

    # ---- config 3 (BASELINE): double-Gauss 1e9-ray dense pupil sweep, one field, on ONE GPU: statistics only
    #      (0 B/ray of output) and with the spot diagram + mask (17 B/ray = 17 GB)
    if len(sys.argv) > 3 and sys.argv[3] == "config3":
        torch.cuda.empty_cache()
        Pd = ort.prescriptions.DOUBLE_GAUSS
        sd = ort.solve(Pd["surfaces"], Pd["a"], Pd["h"])
        pd_ = ort.host._full_trace_setup(sd.layout, sd, [0.7], 64, None, ctx)
        ctx.set_layout(pd_["ext"], pd_["K"])
        ny3, nx3 = 44722, 22361
        ys3 = torch.from_numpy(np.linspace(pd_["y1"][0], pd_["y2"][0], ny3)).to(dev)
        xs3 = torch.from_numpy(np.linspace(0.0, pd_["y_EP"], nx3)).to(dev)
        st3 = torch.zeros(ort.STATS_BYTES, dtype=torch.uint8, device=dev)
        fld3 = dict(u=float(pd_["u"][0]), h_prime=float(pd_["h_prime"][0]))
        NN3 = ny3 * nx3
        ms, best = timed(ctx, lambda: ctx.trace3d_grid_dev([fld3], ys3.data_ptr(), ny3, xs3.data_ptr(), nx3, pd_["stop"], pd_["a_stop"],
                                                          dict(stats=st3.data_ptr()), stream=st), reps=3, warm=1)
        rec = np.frombuffer(st3.cpu().numpy().tobytes(), dtype=ort.STATS_DTYPE)[0]
        out["config3_1e9_rays_stats_only"] = {"rays": NN3, "ms": ms, "rays_per_s": NN3 / ms * 1e3, "intersections_per_s_x10": NN3 * 10 / ms * 1e3,
                                              "fp64_frac": NN3 * 723 / ms / 1e9 / peak, "kept": int(rec["n_kept"]), "rms_mm": ort.rms_from_stats(rec),
                                              "n_strict": int(rec["n_strict"])}
        ex3 = torch.empty(NN3, dtype=torch.float64, device=dev); ey3 = torch.empty(NN3, dtype=torch.float64, device=dev)
        mk3 = torch.empty(NN3, dtype=torch.uint8, device=dev)
        ms, best = timed(ctx, lambda: ctx.trace3d_grid_dev([fld3], ys3.data_ptr(), ny3, xs3.data_ptr(), nx3, pd_["stop"], pd_["a_stop"],
                                                          dict(ex=ex3.data_ptr(), ey=ey3.data_ptr(), mask=mk3.data_ptr(), stats=st3.data_ptr()),
                                                          stream=st), reps=3, warm=1)
        out["config3_1e9_rays_spot_and_mask"] = {"rays": NN3, "ms": ms, "rays_per_s": NN3 / ms * 1e3, "GBps": NN3 * 17 / ms / 1e6,
                                                 "fp64_frac": NN3 * 723 / ms / 1e9 / peak, "mask_sum_equals_kept": int(mk3.sum()) == int(rec["n_kept"])}
        del ex3, ey3, mk3

    res = d_o.cpu().numpy()
    out["candidates_rms_range"] = [float(np.nanmin(res[:, 3])), float(np.nanmax(res[:, 3]))]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
