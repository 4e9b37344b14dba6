"""Host-side mirror of the reference API (ort_b200.host) exercised WITHOUT a GPU: the same host code
with an oracle-backed backend injected (tests/oracle_backend.py).  Checks the reference's known
answers and that the product host logic agrees with the independent oracle prelude."""
import json
import math
import os

import numpy as np
import pytest

from oracle_backend import OracleBackend

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_known_answers.json")) as f:
    KAT = json.load(f)


@pytest.fixture(scope="module")
def be(orc):
    return OracleBackend()


@pytest.fixture(scope="module")
def cooke(ort, be):
    P = ort.prescriptions.COOKE
    return ort.solve(P["surfaces"], P["a"], P["h"], backend=be)


def test_solve_known_answers(ort, cooke):
    C = KAT["cooke_triplet"]
    assert abs(cooke.f - C["EFL"]) < 1e-3 and abs(cooke.EBFD - C["BFL"]) < 1e-3 and cooke.stop == 5
    assert abs(cooke.N - 1 / (2 * C["NA"])) < 1e-3 and abs(cooke.FOV - 2 * C["HFOV"]) < 1e-3
    yui = np.array(C["yui"])
    assert np.max(np.abs(cooke.marginal.y - yui[:, 0])) < 1e-2 and np.max(np.abs(cooke.marginal.u - yui[:, 1])) < 1e-2
    fl = ort.flatten(cooke)
    assert math.isclose(fl["f"], cooke.f, rel_tol=1e-12) and math.isclose(fl["EBFD"], cooke.EBFD, rel_tol=1e-12)
    assert math.isclose(fl["EFFD"], cooke.EFFD, rel_tol=1e-12)
    assert math.isclose(fl["P1"], cooke.P1, rel_tol=1e-9) and math.isclose(fl["P2"], cooke.P2, rel_tol=1e-9)


def test_host_prelude_equals_oracle_prelude(ort, pre, be, cooke):
    P = ort.prescriptions.COOKE
    so = pre.solve(P["surfaces"], P["a"], P["h"])
    assert cooke.f == so.f and cooke.EBFD == so.EBFD and cooke.stop == so.stop
    assert np.array_equal(cooke.marginal.yu, so.marginal.yu) and np.array_equal(cooke.chief.z, so.chief.z)
    assert np.array_equal(cooke.M, so.M)
    rm, rmo = ort.trace_marginal_ray(P["surfaces"], cooke, backend=be), pre.trace_marginal_ray_real(P["surfaces"], so)
    assert np.array_equal(rm.y, rmo.y) and np.array_equal(rm.z, rmo.z)
    rc, rco = ort.trace_chief_ray(P["surfaces"], cooke, backend=be), pre.trace_chief_ray_real(P["surfaces"], so)
    assert np.array_equal(rc.y, rco.y) and np.array_equal(rc.z, rco.z)
    for H in (0.0, 0.7, 1.0):
        e, eo = ort.full_trace(cooke, H, backend=be), pre.full_trace(so, H)
        assert len(e.x) == len(eo.x)
        assert np.array_equal(e.x, eo.x) and np.array_equal(e.y, eo.y) and np.array_equal(e.t, eo.t)
        assert np.allclose(e.r, eo.r, rtol=1e-15) and math.isclose(e.RMS, eo.RMS, rel_tol=1e-12)
        assert e.nu == so.marginal.nu[-1] and e.H == H
        wx, wy = ort.wavegrad(e)
        assert np.array_equal(wx, e.x * e.nu / ort.LAMBDA)


def test_full_trace_singlet_rms(ort, be):
    SG = KAT["singlet"]
    P = ort.prescriptions.SINGLET
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=be)
    for H, rms in zip(SG["H"], SG["RMS"]):
        assert abs(ort.full_trace(s, H, backend=be).RMS - rms) < SG["spot_scale_atol"]
    with pytest.raises(ValueError):
        ort.full_trace(s, 1.5, backend=be)                   # DomainError: |H| <= 1 (PupilSampling.jl:89)


def test_raytrace_dispatch(ort, be, cooke):
    P = ort.prescriptions.COOKE
    S = P["surfaces"]
    ray = ort.raytrace(S, 1.0, 0.0, ort.RealRay, backend=be)                    # 2-D meridional
    par = ort.raytrace(S, 1.0, 0.0, backend=be)                                 # paraxial, matrix form
    assert len(ray.y) == len(S) and abs(ray.y[-1] - par.y[-1]) < 1e-2
    xv, yv = ort.raytrace(S, 1.0, 0.0, 0.0, 0.0, ort.VectorRealRay, backend=be)  # 3-D skew
    assert np.allclose(yv, ray.y[1:], rtol=1e-13)
    ya, wa, ci = ort.raytrace(cooke.lens, np.array([1.0, 2.0]), np.array([0.0, 0.0]), backend=be)
    assert ya.shape == (len(cooke.lens.tau) + 1, 2) and np.allclose(ya[:, 1], 2 * ya[:, 0])
    # clip (test/runtests.jl:252-257)
    ybar = np.abs(cooke.chief.y[1:-1])
    with np.errstate(divide="ignore"):
        slope = abs(cooke.chief.u[0] * np.min(P["a"] / ybar))
    y0 = -slope * cooke.EP.t
    r1 = ort.raytrace(cooke.lens, y0, slope, P["a"], clip=True, backend=be)
    r2 = ort.raytrace(cooke.lens, y0 - 1e-12, slope, P["a"], clip=True, backend=be)
    assert not np.isnan(r1.ynu).any() and np.isnan(r2.ynu).any()


def test_raybasis_lagrange_invariant(ort, be, cooke):
    """test/runtests.jl:94-127"""
    s = -cooke.f + cooke.EFFD
    rb = ort.raytrace(cooke, 10.0, s)
    m, c = rb.marginal, rb.chief
    assert m.y[-1] == 0.0
    Hv = c.nu * m.y - m.nu * c.y
    assert np.allclose(Hv, rb.H, rtol=1e-9)
    assert math.isclose(c.y[-1], -10.0, rel_tol=1e-9)
    e = ort.full_trace(cooke.layout, rb, 32, backend=be)      # RayBasis grid mode (PupilSampling.jl:124-127)
    assert e.H == 1.0 and len(e.x) > 0 and math.isfinite(e.RMS)


def test_transfer_and_reverse(ort, be, cooke):
    v = np.array([3.0, -0.05])
    w = ort.transfer(cooke, v, -40.0, 12.0, backend=be)
    assert np.allclose(ort.reverse_transfer(cooke.M, w, 12.0, -40.0, backend=be), v, rtol=1e-11)
    rv = ort.reverse_transfer(cooke.M, [1.0, 0.0], 0.0, 0.0, backend=be)
    assert math.isclose(-rv[0] / rv[1], cooke.EFFD, rel_tol=1e-12)       # test/runtests.jl:238
    batch = ort.transfer(cooke.M, np.tile(v, (5, 1)), 0.0, 0.0, backend=be)
    assert batch.shape == (5, 2) and np.allclose(batch[0], cooke.M @ v)


def test_tsa_and_sa(ort, be, cooke):
    y, eps_ = ort.TSA(ort.prescriptions.COOKE["surfaces"], cooke, backend=be)
    B1 = ort.SA(y, eps_, 9)[0]
    assert abs(B1 / KAT["cooke_triplet"]["W040_book"] - 1) < 0.05        # test/runtests.jl:277-278
    with pytest.raises(ValueError):
        ort.SA(y, eps_, 4)


def test_layout_rejects_polynomials(ort):
    with pytest.raises(ValueError):
        ort.Layout(ort.prescriptions.COOKE["surfaces"], p=[lambda y: y ** 4])      # closures cannot cross the C ABI
    L = ort.Layout(ort.prescriptions.PARABOLA["surfaces"])
    assert L.aspheric and L.K[1] == -1.0 and L.P is None


def test_layout_polynomials_in_coefficient_form(ort, orc, be):
    """EXTENSION: p given as coefficient sequences.  Zero coefficients == no polynomial (bit for bit); a y^4 term on
    the first surface of a singlet moves the marginal focus the way a weaker rim does; 3-D == 2-D on meridional rays."""
    S = ort.prescriptions.SINGLET["surfaces"]
    L0 = ort.Layout(S)
    Lz = ort.Layout(S, p=[None, [0.0, 0.0, 0.0, 0.0, 0.0], None])
    assert Lz.P is None and not Lz.aspheric
    Lp = ort.Layout(S, p=[None, [0.0, 0.0, 0.0, 0.0, -2e-7], None])
    assert Lp.P.shape == (3, 5) and Lp.aspheric
    y, U = 12.0, 0.0
    r0 = ort.raytrace(L0, y, U, ort.RealRay, backend=be)
    rp = ort.raytrace(Lp, y, U, ort.RealRay, backend=be)
    # p(y) = -2e-7 y^4 flattens the rim: sag -4.1e-3 mm and surface slope -1.4e-3 at y = 12 -> a weaker refraction
    assert rp.u[1] != r0.u[1] and abs(rp.u[1]) < abs(r0.u[1]) and abs(rp.u[1] - r0.u[1]) < 2e-3
    assert rp.y[2] != r0.y[2]
    xv, yv = ort.raytrace(Lp, y, 0.0, U, 0.0, ort.VectorRealRay, backend=be)
    assert np.allclose(yv, rp.y[1:], rtol=1e-12, atol=1e-12) and np.all(xv == 0.0)
    # complex-step derivative of the restatement == analytic derivative to rounding
    orc.set_poly(Lp.P)
    rt, ts, fl = orc.trace2d(S, y, U, K=np.zeros(3), aspheric=True)
    orc.set_poly(None)
    assert np.array_equal(rt[:, 0], rp.y)


def test_merge_stats_matches_direct(ort):
    rng = np.random.default_rng(3)
    ex, ey, r = rng.normal(0.3, 0.05, 5000), rng.normal(-0.1, 0.02, 5000), rng.uniform(0, 10, 5000)
    recs = np.zeros(4, dtype=ort.STATS_DTYPE)
    cuts = [0, 1000, 1000, 3500, 5000]                      # includes an empty shard
    for j in range(4):
        sl = slice(cuts[j], cuts[j + 1])
        n = cuts[j + 1] - cuts[j]
        recs[j]["n_kept"] = n
        if n:
            recs[j]["mean_x"], recs[j]["mean_y"] = ex[sl].mean(), ey[sl].mean()
            recs[j]["m2_x"], recs[j]["m2_y"] = ((ex[sl] - ex[sl].mean()) ** 2).sum(), ((ey[sl] - ey[sl].mean()) ** 2).sum()
            recs[j]["r_max"] = r[sl].max()
        recs[j]["n_clip"] = j
    m = ort.merge_stats(recs)
    assert m["n_kept"] == 5000 and m["n_clip"] == 6 and m["r_max"] == r.max()
    assert math.isclose(m["mean_x"], ex.mean(), rel_tol=1e-13) and math.isclose(m["mean_y"], ey.mean(), rel_tol=1e-13)
    assert math.isclose(m["m2_x"], ((ex - ex.mean()) ** 2).sum(), rel_tol=1e-12)
    x2, y2 = np.concatenate([ex, -ex]), np.concatenate([ey, ey])
    direct = math.sqrt((((x2 - x2.mean()) ** 2).sum() + ((y2 - y2.mean()) ** 2).sum()) / len(x2))
    assert math.isclose(ort.rms_from_stats(m), direct, rel_tol=1e-12)


def test_shard_rows(ort):
    for ny, w in ((10, 3), (5792, 8), (7, 8), (0, 2)):
        spans = [ort.distributed.shard_rows(ny, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == ny
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_aberrations_and_ray_error_functors(ort, be, cooke):
    """test/runtests.jl:205-229 (coefficients) and :290-310 (transverse ray error functors)"""
    SE = KAT["cooke_triplet"]["seidel"]
    P = ort.prescriptions.COOKE
    ab = ort.aberrations(P["surfaces"], cooke, dn=SE["dn"], backend=be)
    alpha = 2 * cooke.marginal.u[-1] / ort.LAMBDA
    for name, div in SE["total_divisors"].items():
        assert abs(getattr(ab, name) - alpha * SE[name] / div) < 0.25, name
    assert np.allclose(ab.sagittal, ab.petzval + ab.astigmatism / 2) and np.allclose(ab.tangential, ab.petzval + 1.5 * ab.astigmatism)
    ab0 = ort.aberrations(P["surfaces"], cooke, backend=be)                     # no dispersion
    W040, W131, W222, W220P, W311 = (SE[k] for k in ("W040", "W131", "W222", "W220P", "W311"))
    ex, ey = ab0.ray_error(0.0, 1.0, 1.0)
    assert abs(ey - (W040 + 3 * W131 + 3 * W222 + W220P + W311)) < 1e-3          # :294
    ex, ey = ab0.ray_error(1.0, 0.0, 1.0)
    assert abs(ex - (W040 + W222 + W220P)) < 1e-3                                # :295
    rng = np.random.default_rng(5)
    rho, th, H = rng.uniform(), 2 * math.pi * rng.uniform(), rng.uniform()
    x, y = rho * math.sin(th), rho * math.cos(th)
    ex, ey = ab0.ray_error(x, y, H)
    assert abs(ey - (W040 * rho ** 3 * math.cos(th) + W131 * rho ** 2 * H * (2 + math.cos(2 * th)) +
                     (3 * W222 + W220P) * rho * H ** 2 * math.cos(th) + W311 * H ** 3)) < 1e-3       # :303-306
    assert abs(ex - (W040 * rho ** 3 * math.sin(th) + W131 * rho ** 2 * H * math.sin(2 * th) +
                     (W222 + W220P) * rho * H ** 2 * math.sin(th))) < 1e-3                           # :307-309
    assert abs(ab0(1.0, 0.0, 0.0) - ab0.W040) < 1e-12 and ab0(0.0, 0.3, 0.5) == 0.0
    with pytest.raises(ValueError):
        ab0(1.5, 0.0, 0.0)
    merit, table = ort.seidel_merit(be, ort.prescriptions.perturbed_triplets(16), P["a"], P["h"])
    assert merit.shape == (16,) and np.all(merit > 0) and table.shape == (16, 16)


def test_full_trace_candidates_host_logic(ort, be, cooke):
    """population form == the single-system full_trace, candidate by candidate (oracle backend)"""
    import math
    P = ort.prescriptions.COOKE
    a, h = P["a"], P["h"]
    RtnK = ort.prescriptions.perturbed_triplets(3)
    spot, aim = ort.full_trace_candidates(RtnK, a, h, 0.7, 32, backend=be)
    assert spot.shape == (3, 4) and aim.shape == (3, 24)
    for c in range(3):
        Sc = RtnK[c, :3].T.copy()
        system = ort.solve(Sc, a, h, backend=be)
        e = ort.full_trace(Sc, system, 0.7, 32, backend=be)
        assert spot[c, 0] * 2 == len(e.x)
        assert abs(math.sqrt(spot[c, 3] ** 2 + spot[c, 1] ** 2) / e.RMS - 1) < 1e-9
    with pytest.raises(ValueError):
        ort.full_trace_candidates(RtnK, a, h, 1.2, backend=be)


def test_vignetting_host_equals_oracle(ort, pre, be, cooke):
    """vignetting(system, a): the host mirror == the CPU restatement (src/Vignetting.jl:1-30), incl. the index lists"""
    P = ort.prescriptions.COOKE
    so = pre.solve(P["surfaces"], P["a"], P["h"])
    for a in (P["a"], np.asarray(P["a"]) * 0.9, np.asarray(P["a"]) * 2.5, [14.7, 14.7, 10.8, 3.0, 10.3, 11.6, 30.0]):
        h, o = ort.vignetting(cooke, a, backend=be), pre.vignetting(so, a)
        assert np.array_equal(h.M, o.M, equal_nan=True) and np.array_equal(h.FOV, o.FOV, equal_nan=True)
        assert h.un == o.un
        for k in ("limit", "partial", "full"):
            assert list(getattr(h, k)) == list(getattr(o, k))
    assert list(ort.vignetting(cooke, backend=be).partial) == [1, 2, 3, 6, 7]  # test/runtests.jl:243


def test_sphere_root_forms_error_study(tmp_path):
    """The SIMPLE kernels take the ray-sphere path parameter division-free, s = (G - sqrt(disc)) / (c n1^2) (DESIGN.md
    section 4).  tools/sphere_root_forms.c traces the bench lens in double with both forms and in 80-bit arithmetic: the
    division-free form must stay as accurate as the one with the division at the lens's own radii, and its error must
    grow no faster than ~ eps |R| for a weak surface -- the bound behind the |R| <= 64 L admission rule."""
    import re
    import subprocess
    exe = tmp_path / "srf"
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sphere_root_forms.c")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), src, "-lm"])
    out = subprocess.run([str(exe), "1500"], capture_output=True, text=True, check=True).stdout
    rows = {}
    for line in out.splitlines():
        m = re.match(r"(.*?)\s+stable (\S+)\s+cheap (\S+)\s+centre (\S+)", line)
        rows[m.group(1).strip()] = tuple(float(m.group(i)) for i in (2, 3, 4))
    nominal = rows["double-Gauss, nominal radii"]
    assert nominal[0] < 1e-14 and nominal[1] < 1e-14
    assert rows["surface 2 with R = 1e+04"][1] < 2e-14           # 64 L = 8.8e3 mm for this lens
    assert rows["surface 2 with R = 1e+06"][1] < 2e-12           # ~ eps |R|: two decades of R, two decades of error
    assert rows["surface 2 with R = 1e+06"][0] < 1e-14           # the form with the division does not care


def test_jl_range_is_the_correctly_rounded_interpolation(ort, pre):
    """collect(range(a, b, n)) (src/PupilSampling.jl:121-122) restated from Julia Base's TwicePrecision algorithm in the
    host mirror (vectorised) and, separately, in the oracle prelude (scalar loop): both give identical bits, hit both end
    points exactly and stay within 1 ulp of -- almost always equal to -- the exactly interpolated value.  numpy.linspace
    does not (it is start + i*step in plain Float64)."""
    from fractions import Fraction
    rng = np.random.default_rng(5)
    total = off = 0
    for trial in range(60):
        a, b, n = float(rng.uniform(-40, 40)), float(rng.uniform(-40, 40)), int(rng.integers(2, 1500))
        if trial % 4 == 0:
            a = 0.0                                   # xs = range(0, y_EP, k / 2)
        r = ort.host.jl_range(a, b, n)
        assert np.array_equal(r, pre.jl_range(a, b, n))
        assert r[0] == a and r[-1] == b and len(r) == n
        fa, fb = Fraction(a), Fraction(b)
        exact = np.array([float(fa + (fb - fa) * Fraction(i, n - 1)) for i in range(n)])
        ulp = np.abs(r - exact) / np.spacing(np.maximum(np.abs(exact), 1e-300))
        assert ulp.max() <= 1.0
        total += n
        off += int((ulp > 0).sum())
    assert off <= total // 2000                       # a handful of 1-ulp cases, as in Julia itself
    assert np.array_equal(ort.host.jl_range(0.0, 1.0, 11), np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]))
    assert np.array_equal(ort.host.jl_range(-1.0, 1.0, 5), np.array([-1.0, -0.5, 0.0, 0.5, 1.0]))
    assert np.array_equal(ort.host.jl_range(2.5, 2.5, 3), np.full(3, 2.5))


def test_comm_helpers_without_a_gpu(ort):
    """ort_comm_range is the block partition of distributed.shard_rows; the communicator id comes from NCCL through
    dlopen (no GPU needed); compute still fails loudly"""
    for total, world in ((64, 4), (44722, 8), (5, 8), (0, 3), (65536, 7)):
        got = [ort._lib.comm_range(total, r, world) for r in range(world)]
        assert got == [ort.distributed.shard_rows(total, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == total and all(got[i][1] == got[i + 1][0] for i in range(world - 1))
    ident = ort._lib.comm_unique_id()
    assert len(ident) == ort._lib.COMM_ID_BYTES == 128 and ident != ort._lib.comm_unique_id()
    L = ort._lib.load()
    assert L.ort_comm_init_rank(None, ident, 0, 1) == ort._lib.ORT_EINVAL
    assert L.ort_comm_free(None) == ort._lib.ORT_EINVAL
