"""GPU parity of the 2-D meridional tracer (K2), paraxial y-nu trace (K3) and transfer-matrix
apply (K4) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

from util import bits_equal, n_bits_differ, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("name,aspheric", [("COOKE", False), ("DOUBLE_GAUSS", False), ("REFLECTIVE", False),
                                           ("PARABOLA", True), ("COOKE", True)])
def test_trace2d_batch(ctx, orc, ort, name, aspheric):
    P = getattr(ort.prescriptions, name)
    S = P["surfaces"]
    K = S[:, 3] if S.shape[1] > 3 else None
    amax = float(P["a"][0])
    rng = np.random.default_rng(5)
    N = 4000
    y0 = rng.uniform(-1.2 * amax, 1.2 * amax, N)
    U0 = rng.uniform(-0.25, 0.25, N)
    y0[:3] = [0.0, amax, -amax]; U0[:3] = 0.0
    yo, Uo, tso, fo = orc.trace2d_batch(S[:, :3], y0, U0, K=K, aspheric=aspheric)
    ctx.set_layout(S[:, :3], K)
    yg, Ug, tsg, fg = ctx.trace2d_batch(y0, U0, aspheric=aspheric)
    # libm (tan/sin/asin/atan) differs by <= 2 ulp between CUDA and glibc (and Julia): a decision
    # sitting exactly on a branch (TIR / miss) may flip for a handful of rays -> compare where flags agree
    same = fg == fo
    assert same.mean() > 0.999
    scale = np.maximum(np.nanmax(np.abs(yo), axis=0), 1.0)
    for a, b, s in ((yg, yo, scale), (Ug, Uo, 1.0), (tsg, tso, np.maximum(np.nanmax(np.abs(tso), axis=0), 1.0))):
        a, b = a[:, same], b[:, same]
        ss = s[same] if isinstance(s, np.ndarray) else s
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.nanmax(np.abs(a - b) / ss) < TOL


def test_parabola_focus_exact(ctx, ort):
    """test/runtests.jl:334-344: a parabolic mirror has zero spherical aberration; the marginal
    real ray crosses the axis at exactly -50.0."""
    S = ort.prescriptions.PARABOLA["surfaces"]
    ctx.set_layout(S[:, :3], S[:, 3])
    y, U, ts, f = ctx.trace2d_batch(np.array([30.0, 10.0, 1.0]), np.zeros(3), aspheric=True)
    z = np.cumsum(ts, axis=0)
    zf = z[-2] - y[-1] / np.tan(U[-1])
    assert np.max(np.abs(zf + 50.0)) < 1e-12


def test_paraxial_batch(ctx, orc, ort):
    S = ort.prescriptions.zoom20()
    tau, phi, n = orc.lens(S)
    assert len(tau) == 40
    rng = np.random.default_rng(42)
    N = 100_003
    y0, w0 = rng.uniform(-10, 10, N), rng.uniform(-0.2, 0.2, N)
    a = np.full(len(tau), 40.0)
    for clip in (False, True):
        yo, wo, co = orc.paraxial_batch(tau, phi, y0, w0, a=a, clip=clip)
        y, w, c = ctx.paraxial_batch(tau, phi, y0, w0, a=a, clip=clip, arith=ort.STRICT)
        assert np.array_equal(c, co)
        assert n_bits_differ(y, yo) == 0 and n_bits_differ(w, wo) == 0
        yf, wf, cf = ctx.paraxial_batch(tau, phi, y0, w0, a=a, clip=clip, arith=ort.FAST)
        agree = cf == co                      # FMA rounding can flip a ray sitting on the 1e-13 clip edge
        assert agree.mean() > 0.9999
        sc = np.maximum(np.abs(y0) + 400 * np.abs(w0), 1.0)
        ok = agree & ~np.isnan(yo)
        assert np.max(np.abs(yf[ok] - yo[ok]) / sc[ok]) < TOL
        assert np.max(np.abs(wf[ok] - wo[ok])) < TOL
    # full table == the reference's rt matrix, incl. NaN fill after the clip row
    y, w, c, ya, wa = ctx.paraxial_batch(tau, phi, y0[:500], w0[:500], a=a, clip=True, arith=ort.STRICT, table=True)
    for j in range(0, 500, 37):
        rt, ci = orc.paraxial_trace(tau, phi, y0[j], w0[j], a=a, clip=True)
        assert ci == c[j]
        assert bits_equal(ya[:, j], rt[:, 0]) and bits_equal(wa[:, j], rt[:, 1])


def test_paraxial_clip_classifier_adversarial(ctx, orc, ort):
    """the FP32 high-word classifier of k_paraxial's clip test (src/RayTracing.jl:135) must agree with the exact FP64 test
    everywhere: rays that hug an aperture to the last ulp, the 1e-13 threshold itself, apertures of every magnitude
    (tiny, huge, +Inf, NaN, 0, negative), more than 32 rows (second mask chunk), NaN / Inf rays.  STRICT bit-exact."""
    rng = np.random.default_rng(11)
    # (1) a lens that leaves y untouched (tau = 0, phi = 0): y_row == y0, so y0 can be placed at will against a[row]
    k = 70
    tau, phi = np.zeros(k), np.zeros(k)
    a = np.full(k, np.inf)
    a[41] = 7.25
    base = 7.25
    offs = np.concatenate([np.arange(-40, 41) * np.spacing(base), [1e-13, 1e-13 + np.spacing(base), 1e-13 - np.spacing(base),
                                                                  2e-13, 5e-14, -1e-13, 3e-7, -3e-7, 1e-6, -1e-6]])
    y0 = np.concatenate([base + offs, -(base + offs), [np.nan, np.inf, -np.inf, 0.0, 1e300, -1e300, 5e-324]])
    w0 = np.zeros_like(y0)
    for arith in (ort.STRICT, ort.FAST):
        yo, wo, co = orc.paraxial_batch(tau, phi, y0, w0, a=a, clip=True)
        y, w, c = ctx.paraxial_batch(tau, phi, y0, w0, a=a, clip=True, arith=arith)
        assert np.array_equal(c, co), np.flatnonzero(c != co)
        assert n_bits_differ(y, yo) == 0 and n_bits_differ(w, wo) == 0
    assert set(np.unique(co)) == {0, 42}
    # (2) apertures of every magnitude and kind, rays of every magnitude, real zoom rows
    S = ort.prescriptions.zoom20()
    tau, phi, n = orc.lens(S)
    N = 60_000
    for trial, amag in enumerate((1e-9, 1e-7, 2.2e-7, 1e-3, 1.0, 25.0, 1e6, 1e30)):
        a = amag * rng.uniform(0.5, 2.0, len(tau))
        a[rng.integers(0, len(a), 4)] = [np.inf, np.nan, 0.0, -1.0][trial % 4]
        y0 = amag * rng.uniform(-3, 3, N)
        w0 = amag * rng.uniform(-0.05, 0.05, N)
        yo, wo, co = orc.paraxial_batch(tau, phi, y0, w0, a=a, clip=True)
        y, w, c = ctx.paraxial_batch(tau, phi, y0, w0, a=a, clip=True, arith=ort.STRICT)
        assert np.array_equal(c, co), (amag, int((c != co).sum()))
        assert n_bits_differ(y, yo) == 0 and n_bits_differ(w, wo) == 0
        cf = ctx.paraxial_batch(tau, phi, y0, w0, a=a, clip=True, arith=ort.FAST)[2]
        assert (cf == co).mean() > 0.9995                     # FMA rounding may move a ray across the 1e-13 edge
    # (3) rays aimed at an aperture edge of a real row: bisect y0 so that |y_row| lands within an ulp of a[row]
    a = np.full(len(tau), 1e9)
    row = 17
    a[row] = 12.5
    lo, hi = np.full(64, 0.0), np.full(64, 60.0)
    w0 = rng.uniform(-0.02, 0.02, 64)
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        # height at `row` from the oracle's table
        hts = np.array([abs(orc.paraxial_trace(tau, phi, m, ww)[0][row + 1, 0]) for m, ww in zip(mid, w0)])
        big = hts - a[row] > 1e-13                      # the reference's own test (:135)
        hi = np.where(big, mid, hi); lo = np.where(big, lo, mid)
    y0 = np.concatenate([lo, hi, np.nextafter(lo, -1), np.nextafter(hi, 100)])
    w0 = np.tile(w0, 4)
    yo, wo, co = orc.paraxial_batch(tau, phi, y0, w0, a=a, clip=True)
    y, w, c = ctx.paraxial_batch(tau, phi, y0, w0, a=a, clip=True, arith=ort.STRICT)
    assert np.array_equal(c, co) and n_bits_differ(y, yo) == 0
    assert 0 < (co == row + 1).sum() < len(co)


def test_paraxial_clip_threshold(ctx, orc, pre, ort):
    """test/runtests.jl:252-257: the half-vignetted chief ray passes; lowered by 1e-12 it is clipped."""
    P = ort.prescriptions.COOKE
    s = pre.solve(P["surfaces"], P["a"], P["h"])
    a = P["a"]
    ybar = np.abs(s.chief.y[1:-1])
    slope = abs(s.chief.u[0] * np.min(a / ybar))
    y = -slope * s.EP.t
    yv, wv, c = ctx.paraxial_batch(s.tau, s.phi, np.array([y, y - 1e-12]), np.array([slope, slope]), a=a,
                                   clip=True, arith=ort.STRICT)
    assert c[0] == 0 and not np.isnan(yv[0])
    assert c[1] != 0 and np.isnan(yv[1])


def test_transfer_batch(ctx, orc, pre, ort):
    P = ort.prescriptions.COOKE
    s = pre.solve(P["surfaces"], P["a"], P["h"])
    rng = np.random.default_rng(3)
    N = 200_001
    v = np.column_stack([rng.uniform(-20, 20, N), rng.uniform(-0.3, 0.3, N)])
    for tau, taup in ((0.0, 0.0), (-123.4, 77.4), (50.0, 0.0)):
        ref = orc.transfer_batch(s.M, tau, taup, v)
        out = ctx.transfer_batch(s.M, tau, taup, v)
        assert n_bits_differ(out, ref) == 0
        refr = orc.transfer_batch(s.M, tau, taup, v, reverse=True)
        outr = ctx.transfer_batch(s.M, tau, taup, v, reverse=True)
        assert n_bits_differ(outr, refr) == 0
        # round trip: reverse(forward(v)) == v  (size-independent property)
        back = ctx.transfer_batch(s.M, tau, taup, out, reverse=True)
        assert np.max(np.abs(back - v)) < 1e-9
    # matrix == surface-by-surface paraxial trace (test/runtests.jl:133-145)
    y, w, _ = ctx.paraxial_batch(s.tau, s.phi, v[:1000, 0], v[:1000, 1], arith=ort.STRICT)
    out = ctx.transfer_batch(s.M, 0.0, 0.0, v[:1000])
    assert np.allclose(out[:, 0], y, rtol=1e-12, atol=1e-12) and np.allclose(out[:, 1], w, rtol=1e-12, atol=1e-12)


def test_error_paths(ctx, ort):
    with pytest.raises(ort.OrtError):
        ctx.set_layout(np.zeros((1, 3)))                      # rows < 2
    ctx.set_layout(ort.prescriptions.COOKE["surfaces"])
    with pytest.raises(ort.OrtError):
        ctx.trace3d_grid([dict(u=0.0)], np.zeros(4), np.zeros(4), 99, 1.0)   # stop out of range
    with pytest.raises(ort.OrtError):
        ctx.trace3d_grid([dict(u=0.0)] * 40, np.zeros(4), np.zeros(4), 5, 1.0)   # too many fields


def test_exact_division_and_sqrt_with_deferred_slow_path(ctx):
    """k_grid<STRICT> takes its divisions and square roots from xdiv / xsqrt: the CUDA library's own fast paths with the
    range test folded into a per-ray flag (flagged rays are traced again with the intrinsics).  Every result that is kept
    must be the intrinsic's, bit for bit: 2^28 operand pairs of four classes (raw bit patterns, moderate magnitudes,
    every exponent, special values).  Moderate operands never raise the flag."""
    r = ctx.selftest_exact_ops(1 << 28, seed=12345)
    assert r["div_tested"] >= 1 << 28 and r["sqrt_tested"] >= 1 << 28
    assert r["div_mismatch"] == 0 and r["sqrt_mismatch"] == 0, r
    assert 0 < r["div_flagged"] < r["div_tested"] and 0 < r["sqrt_flagged"] < r["sqrt_tested"]     # both paths exercised
    assert r["div_flagged_moderate"] == 0 and r["sqrt_flagged_moderate"] == 0, r
    r2 = ctx.selftest_exact_ops(1 << 24, seed=999)
    assert r2["div_mismatch"] == 0 and r2["sqrt_mismatch"] == 0


def test_fp64_peak(ctx):
    tf, ms = ctx.fp64_peak()
    assert 10.0 < tf < 60.0, tf


def test_seidel_candidates(ctx, orc, ort):
    """K7: first-order solve + Seidel sums per candidate == the CPU restatement, bit for bit"""
    P = ort.prescriptions.COOKE
    dn = [0.0, 0.010450, 0.0, 0.019151, 0.0, 0.0, 0.010450, 0.0]
    C = 4099
    RtnK = ort.prescriptions.perturbed_triplets(C)
    RtnK[0] = np.vstack([P["surfaces"].T, np.zeros(8)])            # candidate 0 = the nominal triplet
    ref = orc.seidel_candidates(RtnK, P["a"], P["h"], dn=dn)
    out, per = ctx.seidel_candidates(RtnK, P["a"], P["h"], dn=dn, per_surface=True)
    assert n_bits_differ(out, ref) == 0
    d, per0 = orc.seidel(P["surfaces"], P["a"], P["h"], dn=dn)
    assert n_bits_differ(per[0], per0) == 0
    assert abs(out[0, 0] - 101.181) < 1e-3 and out[0, 2] == 5 and abs(out[0, 4] - 11.46) < 0.25
    # a prescription solve() cannot use (last thickness nonzero) -> NaN row, not a crash
    bad = RtnK[:2].copy(); bad[1, 1, -1] = 3.0
    o2 = ctx.seidel_candidates(bad, P["a"], P["h"])
    assert np.isnan(o2[1]).all() and not np.isnan(o2[0]).any()


def test_vignetting_candidates(ctx, pre, ort):
    """batched vignetting(system, a) == the CPU restatement per candidate: table and classification bit for bit,
    FOV angles to the last ulp of atan"""
    P = ort.prescriptions.COOKE
    C = 300
    RtnK = ort.prescriptions.perturbed_triplets(C)
    RtnK[0] = np.vstack([P["surfaces"].T, np.zeros(8)])
    a = np.asarray(P["a"])
    for a_vig in (None, a * 0.93, np.array([14.7, 14.7, 10.8, 3.0, 10.3, 11.6, 30.0])):
        r = ctx.vignetting_candidates(RtnK, a, P["h"], a_vig=a_vig)
        for c in list(range(0, C, 37)) + [C - 1]:
            so = pre.solve(RtnK[c, :3].T.copy(), a, P["h"])
            o = pre.vignetting(so, a if a_vig is None else a_vig)
            assert n_bits_differ(r["M"][c], o.M) == 0
            assert n_bits_differ(r["FOV"][c][:, 1:], o.FOV[:, 1:]) == 0
            assert np.max(np.abs(r["FOV"][c][:, 0] / o.FOV[:, 0] - 1)) < 1e-15
            assert bool(r["un"][c]) == o.un and r["stop"][c] == so.stop and r["f"][c] == so.f
            code = r["code"][c]
            for bit, name in ((1, "limit"), (2, "partial"), (4, "full")):
                assert list(np.nonzero(code & bit)[0] + 1) == list(getattr(o, name))
    r = ctx.vignetting_candidates(RtnK[:1], a, P["h"])
    assert list(np.nonzero(r["code"][0] & 2)[0] + 1) == [1, 2, 3, 6, 7]          # test/runtests.jl:243
    # the single-system host entry runs the same kernel
    system = ort.solve(P["surfaces"], a, P["h"], backend=ctx)
    v = ort.vignetting(system, backend=ctx)
    o = pre.vignetting(pre.solve(P["surfaces"], a, P["h"]))
    assert n_bits_differ(v.M, o.M) == 0 and list(v.partial) == [1, 2, 3, 6, 7] and v.un == o.un
    bad = RtnK[:2].copy(); bad[1, 1, -1] = 3.0
    r = ctx.vignetting_candidates(bad, a, P["h"])
    assert np.isnan(r["M"][1]).all() and not np.isnan(r["M"][0][:, :3]).any()


class _NoAim:
    """proxy that hides aim2d so the host runs its own secant loop (one 2-ray batch per step)"""

    def __init__(self, ctx):
        self._c = ctx

    def __getattr__(self, name):
        if name == "aim2d":
            raise AttributeError(name)
        return getattr(self._c, name)


@pytest.mark.parametrize("name", ["COOKE", "DOUBLE_GAUSS", "TESSAR", "SINGLET"])
def test_device_aiming_equals_host_loop(ctx, orc, ort, name):
    """f1: the secant loops of trace_marginal_ray / trace_chief_ray / trace_edge_rays run in one kernel launch and
    reproduce the launch-per-step host iteration bit for bit; the aimed rays hit their targets (checked with the oracle)."""
    P = getattr(ort.prescriptions, name)
    S = P["surfaces"]
    s = ort.solve(S, P["a"], P["h"], backend=ctx)
    host = _NoAim(ctx)
    rm_d, rm_h = ort.trace_marginal_ray(S, s, backend=ctx), ort.trace_marginal_ray(S, s, backend=host)
    assert np.array_equal(rm_d.y, rm_h.y) and np.array_equal(rm_d.z, rm_h.z)
    L = ort.Layout(S)
    rc_d, rc_h = ort.trace_chief_ray(L, s, backend=ctx), ort.trace_chief_ray(L, s, backend=host)
    assert np.array_equal(rc_d.y, rc_h.y) and np.array_equal(rc_d.u, rc_h.u)
    a_stop = abs(s.a[s.stop - 1])
    assert abs(rm_d.y[s.stop] - a_stop) < 1.5e-8 and abs(rc_d.y[s.stop]) < 1.5e-8          # test/runtests.jl:271-280
    U = np.array([0.0, 0.5, 1.0]) * rc_d.u[0]
    y_EP = abs(rm_d.y[0])
    y1, y2 = y_EP - np.tan(U) * rc_d.z[0], -y_EP - np.tan(U) * rc_d.z[0]
    e_d, e_h = ort.trace_edge_rays(S, y1, y2, U, s.stop, a_stop, backend=ctx), ort.trace_edge_rays(S, y1, y2, U, s.stop, a_stop, backend=host)
    assert np.array_equal(e_d[0], e_h[0]) and np.array_equal(e_d[1], e_h[1])
    for yy, uu, tgt in [(e_d[0][j], U[j], a_stop) for j in range(3)] + [(e_d[1][j], U[j], -a_stop) for j in range(3)]:
        rt, _, _ = orc.trace2d(S, float(yy), float(uu))
        assert abs(rt[s.stop, 0] - tgt) < 1e-12 * a_stop
    l0 = ctx.launch_count()
    ort.trace_edge_rays(S, y1, y2, U, s.stop, a_stop, backend=ctx)
    assert ctx.launch_count() - l0 == 1                                                    # one launch for all 6 solves


def test_c_abi_from_plain_c(ctx, ort, tmp_path):
    """the boundary is a C ABI: examples/full_trace.c compiled with gcc, linked against libort_b200.so and run"""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "full_trace")
    libdir = os.path.dirname(ort._lib.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "full_trace.c"), "-o", exe, "-L", libdir, "-lort_b200", "-lm",
                           "-Wl,-rpath," + libdir])
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=ctx)
    p = ort.host._full_trace_setup(s.layout, s, [0.7], 64, None, ctx)
    args = [repr(float(v)) for v in (p["y1"][0], p["y2"][0], p["y_EP"], p["u"][0], p["h_prime"][0], p["focus"])]
    out = subprocess.run([exe] + args, capture_output=True, text=True, check=True).stdout
    kv = dict(item.split("=") for item in out.split())
    e = ort.full_trace(s, 0.7)
    assert int(kv["n_kept"]) == len(e.x) // 2
    assert abs(float(kv["rms"]) - e.RMS) < 1e-12 * 25 and abs(float(kv["ex0"]) - e.x[1]) < 1e-12 * 25


def test_merge_stats_c_equals_python(ctx, ort):
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=ctx)
    p = ort.host._full_trace_setup(s.layout, s, [0.0, 1.0], 64, None, ctx)
    ctx.set_layout(p["ext"], p["K"])
    ys = np.stack([np.linspace(p["y1"][j], p["y2"][j], 64) for j in range(2)])
    xs = np.linspace(0, p["y_EP"], 32)
    flds = [dict(u=float(p["u"][j]), h_prime=float(p["h_prime"][j])) for j in range(2)]
    recs = np.stack([ctx.trace3d_grid(flds, ys[:, lo:hi], xs, p["stop"], p["a_stop"])["stats"]
                     for lo, hi in (ort.distributed.shard_rows(64, r, 4) for r in range(4))])
    mc = ort._lib.merge_stats_c(recs)
    whole = ctx.trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"])["stats"]
    for f in range(2):
        mp = ort.merge_stats(recs[:, f])
        assert mc[f]["n_kept"] == mp["n_kept"] == whole[f]["n_kept"]
        assert mc[f]["mean_y"] == mp["mean_y"] and mc[f]["m2_x"] == mp["m2_x"]
        assert abs(ort._lib.rms_from_stats_c(mc[f:f + 1]) - ort.rms_from_stats(whole[f])) < 1e-13


def test_tsa_fan_on_gpu(ctx, orc, pre, ort):
    """a15: TSA (src/SeidelAberrations.jl:116-135) = one batch of the 2-D kernel; vs the oracle prelude and the
    reference's SA-fit check (test/runtests.jl:277-278)"""
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=ctx)
    so = pre.solve(P["surfaces"], P["a"], P["h"])
    y, e = ort.TSA(P["surfaces"], s, backend=ctx)
    yo, eo = pre.tsa(P["surfaces"], so)
    # the fan is anchored on the real marginal ray, which the reference's secant loop only aims to |dy_stop| <= sqrt(eps)
    # (src/RayTracing.jl:223-233): GPU and CPU libm differ in the last ulp, so the two aimed rays agree to ~1e-8, not 1e-12
    assert np.max(np.abs(y - yo)) < 1e-7 and np.max(np.abs(e - eo)) < 1e-7
    assert abs(ort.SA(y, e, 9)[0] / -0.186575 - 1) < 0.05
