"""The CPU oracle against every known-answer test the reference holds for the hot path
(test/runtests.jl; constants in tests/golden/reference_known_answers.json).  This is what pins the
oracle: Julia is not installed, so the reference itself cannot be run."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_known_answers.json")) as f:
    KAT = json.load(f)


def _mat(rows):
    return np.array([[np.inf if v == "Inf" else float(v) for v in r] for r in rows])


COOKE = KAT["cooke_triplet"]
S = _mat(COOKE["surfaces"])
A = np.array(COOKE["a"])


@pytest.fixture(scope="module")
def system(pre):
    return pre.solve(S, A, COOKE["h_prime"])


def test_system_properties(system):
    """test/runtests.jl:53-60"""
    tol = COOKE["system_scale_atol"]
    assert abs(system.f - COOKE["EFL"]) < tol
    assert abs(system.EBFD - COOKE["BFL"]) < tol
    assert abs(system.N - 1 / (2 * COOKE["NA"])) < tol
    assert abs(system.FOV - 2 * COOKE["HFOV"]) < tol
    assert abs(S[1:, 1].sum() - COOKE["VL"]) < tol
    assert system.stop == COOKE["stop"]


def test_raytrace_validation(system):
    """test/runtests.jl:62-113: marginal / chief y, u at all 9 planes and the incidence angles"""
    tol = COOKE["trace_scale_atol"]
    yui, ybu = np.array(COOKE["yui"]), np.array(COOKE["ybar_ubar_ibar"])
    assert np.max(np.abs(system.marginal.y - yui[:, 0])) < tol
    assert np.max(np.abs(system.marginal.u - yui[:, 1])) < tol
    assert np.max(np.abs(system.chief.y - ybu[:, 0])) < tol
    assert np.max(np.abs(system.chief.u - ybu[:, 1])) < tol
    # incidences(surfaces, system)  src/RayTracing.jl:337-353
    R = S[1:, 0]
    n = system.marginal.n[:len(R)]
    i_m = (system.marginal.nu[:len(R)] + n * system.marginal.y[1:-1] / R) / n
    i_c = (system.chief.nu[:len(R)] + n * system.chief.y[1:-1] / R) / n
    assert np.max(np.abs(i_m - yui[1:-1, 2])) < tol
    assert np.max(np.abs(i_c - ybu[1:-1, 2])) < tol


def test_matrix_equals_trace_and_lagrange(system, orc):
    """test/runtests.jl:115-146: Lagrange invariant; system.M * v == surface-by-surface trace == transfer"""
    m, c = system.marginal, system.chief
    H = c.nu * m.y - m.nu * c.y
    assert np.allclose(H, system.H, rtol=1e-12, atol=1e-12)
    rng = np.random.default_rng(0)
    for _ in range(20):
        y_in, u_in = rng.uniform(-15, 15), rng.uniform(-0.3, 0.3)
        rt, _ = orc.paraxial_trace(system.tau, system.phi, y_in, u_in)
        v = system.M @ np.array([y_in, u_in])
        assert np.allclose(v, rt[-1], rtol=1e-12, atol=1e-12)
        s = rng.uniform(-1000, 0)
        vt = orc.transfer(system.M, [y_in, u_in], -s, 0.0)
        rt2, _ = orc.paraxial_trace(system.tau, system.phi, y_in + u_in * -s, u_in)
        assert np.allclose(vt, rt2[-1], rtol=1e-11, atol=1e-11)


def test_transfer_matrix_flatten_and_reverse(system, orc):
    """test/runtests.jl:231-239"""
    M = system.M
    f = -1.0 / M[1, 0]
    assert math.isclose(f, system.f, rel_tol=1e-12)
    assert math.isclose(M[0, 0] * f, system.EBFD, rel_tol=1e-12)
    assert math.isclose(-M[1, 1] * f, system.EFFD, rel_tol=1e-12)
    assert math.isclose(-M[1, 1] * f + f, system.P1, rel_tol=1e-9)
    assert math.isclose(M[0, 0] * f - f, system.P2, rel_tol=1e-9)
    v = orc.reverse_transfer(M, [1.0, 0.0], 0.0, 0.0)
    assert math.isclose(-v[0] / v[1], system.EFFD, rel_tol=1e-12)
    # reverse_transfer inverts transfer
    w = orc.transfer(M, [3.0, -0.05], -40.0, 12.0)
    back = orc.reverse_transfer(M, w, 12.0, -40.0)
    assert np.allclose(back, [3.0, -0.05], rtol=1e-11, atol=1e-12)


def test_vignetting_clip_threshold(system, orc):
    """test/runtests.jl:241-258: the half-vignetted ray passes, the same ray - 1e-12 is clipped
    (the > 1e-13 rule of src/RayTracing.jl:135), and the partially vignetting surfaces."""
    ybar = np.abs(system.chief.y[1:-1])
    y = np.abs(system.marginal.y[1:-1])
    with np.errstate(divide="ignore"):
        slope = abs(system.chief.u[0] * np.min(A / ybar))           # FOV[2,2] of src/Vignetting.jl:17-24
    y0 = -slope * system.EP.t
    rt_half, c0 = orc.paraxial_trace(system.tau, system.phi, y0, slope, a=A, clip=True)
    rt_clip, c1 = orc.paraxial_trace(system.tau, system.phi, y0 - 1e-12, slope, a=A, clip=True)
    assert c0 == 0 and not np.isnan(rt_half).any()
    assert c1 != 0 and np.isnan(rt_clip).any()
    unvig = (A >= y + ybar) | np.isclose(A, y + ybar)                # src/Vignetting.jl:13
    full = A <= np.where(ybar - y < y, np.nan, ybar - y)
    partial = [i + 1 for i in range(len(A)) if not unvig[i] and not full[i]]
    assert partial == COOKE["vignetting_partial"]


def test_vignetting_analysis(system, pre):
    """test/runtests.jl:241-250 on the restated vignetting(system, a): vig.partial == [1,2,3,6,7]; a system re-solved
    at each of the three limiting image heights has some non-stop surface whose aperture equals (isapprox) the
    matching column of its own table."""
    vig = pre.vignetting(system, A)
    assert list(vig.partial) == COOKE["vignetting_partial"] == [1, 2, 3, 6, 7]
    assert list(vig.limit) == [] and list(vig.full) == [] and vig.un is False
    idx = [i for i in range(len(A)) if i != system.stop - 1]
    for i, hp in enumerate(vig.FOV[:, 2]):
        M = pre.vignetting(pre.solve(S, A, hp)).M
        assert np.any(pre._isapprox(A[idx], M[idx, i + 2]))
    assert abs(vig.FOV[1, 0] / 2 - math.degrees(math.atan(vig.FOV[1, 1]))) < 1e-13
    # an aperture stopped down below the marginal ray is "limit"; a wide-open one is unvignetted
    small = A.copy(); small[0] = 10.0
    assert list(pre.vignetting(system, small).limit) == [1]
    assert pre.vignetting(system, A * 3).un is True


def test_real_raytracing(system, pre, orc):
    """test/runtests.jl:260-286"""
    rt_par, _ = orc.paraxial_trace(system.tau, system.phi, 1.0, 0.0)
    u_par = rt_par[:, 1] / np.append(S[:, 2], S[-1, 2])[:len(rt_par)]
    rt_real, ts, fl = orc.trace2d(S, 1.0, 0.0)
    R, t = S[:, 0], S[:, 1]
    k = len(R) - 1
    eps_th = (1.0 / np.min(np.abs(R))) ** 3 / 6 * k * (k + 1) / 2
    eps_y = eps_th * np.max(t)
    assert np.sum(np.abs(u_par[1:] - rt_real[1:, 1])) < eps_th
    assert np.sum(np.abs(rt_par[1:, 0] - rt_real[1:, 0])) < eps_y
    atol = math.sqrt(np.finfo(float).eps)
    rm = pre.trace_marginal_ray_real(S, system, atol=atol)
    assert abs(rm.y[system.stop] - A[system.stop - 1]) < atol
    rt, _, _ = orc.trace2d(S, rm.y[0], rm.u[0])
    assert np.allclose(rt[1:], rm.yu[1:-1], atol=atol, rtol=0)
    rc = pre.trace_chief_ray_real(S, system, atol=atol)
    assert abs(rc.y[system.stop]) < atol
    # transfer(real_chief, surface_to_focus(EBFD, real_chief, marginal)) ~ chief.y[end]
    tfoc = system.EBFD - (rc.z[-2] - system.marginal.z[-2])
    assert math.isclose(rc.y[-2] + math.tan(rc.u[-2]) * tfoc, system.chief.y[-1], rel_tol=1e-2)
    y_vertex = rc.y[1] - math.tan(rc.u[0]) * rc.z[1]
    rt, _, _ = orc.trace2d(S, y_vertex, rc.u[0])
    assert np.allclose(rt[1:], rc.yu[1:-1], atol=atol, rtol=0)
    # SA(TSA(...)..., 9)[1] within 5 % of the book's W040
    y_xp, eps_ = pre.tsa(S, system)
    yp = y_xp / y_xp.max()
    Am = np.column_stack([yp ** q for q in range(3, 10, 2)])
    B1 = np.linalg.lstsq(Am, eps_, rcond=None)[0][0]
    assert abs(B1 / COOKE["W040_book"] - 1) < COOKE["SA_fit_rtol"]


def test_aspheric_parabola_exact(pre):
    """test/runtests.jl:334-344: real and paraxial focus of a parabolic mirror == -50.0 exactly"""
    P = KAT["parabola"]
    M = _mat(P["surfaces_RtnK"])
    sysm = pre.solve(M[:, :3], P["a"], P["h_prime"], K=M[:, 3])
    rm = pre.trace_marginal_ray_real(M[:, :3], sysm, K=M[:, 3], aspheric=True)
    assert sysm.marginal.z[-1] == P["focus_z_exact"]
    assert rm.z[-1] == P["focus_z_exact"]


def test_full_trace_and_pupil_sampling(system, pre, orc):
    """test/runtests.jl:348-373"""
    rm = pre.trace_marginal_ray_real(S, system)
    rc = pre.trace_chief_ray_real(S, system)
    ub = math.tan(rc.u[0])
    yv = rc.y[1] - ub * rc.z[1]
    _, ym, _, _ = orc.trace3d(S, rm.y[0], 0.0, 0.0, 0.0)
    _, yc, _, _ = orc.trace3d(S, yv, 0.0, ub, 0.0)
    assert math.isclose(rm.y[-2], ym[-1], rel_tol=1e-12)
    assert math.isclose(rc.y[-2], yc[-1], rel_tol=1e-9)
    ext = np.vstack([S, [np.inf, 0.0, 1.0]])
    ext[-2, 1] = system.EBFD
    _, ye, _, _ = orc.trace3d(ext, yv, 0.0, ub, 0.0)
    tfoc = system.EBFD - (rc.z[-2] - system.marginal.z[-2])
    assert math.isclose(ye[-1], rc.y[-2] + math.tan(rc.u[-2]) * tfoc, rel_tol=1e-9)
    SG = KAT["singlet"]
    ss = pre.solve(_mat(SG["surfaces"]), SG["a"], SG["h_prime"])
    for H, rms in zip(SG["H"], SG["RMS"]):
        e = pre.full_trace(ss, H)
        assert abs(e.RMS - rms) < SG["spot_scale_atol"], (H, e.RMS)
        n = len(e.x) // 2
        assert np.array_equal(e.x[:n], -e.x[n:]) and np.array_equal(e.y[:n], e.y[n:])     # mirror :140-141
        assert np.allclose(e.t[n:], math.pi - e.t[:n]) and e.r.max() == 1.0             # :142-144


def test_vector_refraction_reflection(orc):
    """test/runtests.jl:376-387: 3-D == 2-D through a reflective (n < 0) stack"""
    P = KAT["reflective_stack"]
    M = _mat(P["surfaces"])
    rt, _, _ = orc.trace2d(M, P["y0"], 0.0)
    _, yv, _, _ = orc.trace3d(M, P["y0"], 0.0, 0.0, 0.0)
    assert math.isclose(rt[-1, 0], yv[-1], rel_tol=1e-12)


def test_hypot_is_correctly_rounded(orc):
    """Julia's hypot (fma branch) is correctly rounded; check against exact integer arithmetic."""
    from fractions import Fraction
    rng = np.random.default_rng(1)
    for _ in range(2000):
        x, y = rng.uniform(-20, 20), rng.uniform(-20, 20)
        h = orc.hypot(x, y)
        exact = Fraction(x) ** 2 + Fraction(y) ** 2
        lo, hi = np.nextafter(h, 0), np.nextafter(h, np.inf)
        # h is the closest double to sqrt(exact): compare squared midpoints
        assert ((Fraction(lo) + Fraction(h)) / 2) ** 2 <= exact <= ((Fraction(h) + Fraction(hi)) / 2) ** 2


def test_pairwise_sum_and_sigma(orc):
    rng = np.random.default_rng(2)
    a = rng.normal(size=5000)
    assert math.isclose(orc.pairwise_sum(a), math.fsum(a), rel_tol=1e-13, abs_tol=1e-12)
    ex, ey = rng.normal(size=3000), rng.normal(size=3000) + 2
    ref = math.sqrt(((ex - ex.mean()) ** 2).sum() / 3000 + ((ey - ey.mean()) ** 2).sum() / 3000)
    assert math.isclose(orc.sigma(ex, ey), ref, rel_tol=1e-13)


def test_oracle_edge_cases(orc):
    # miss -> NaN positions from that surface on, flagged; k untouched afterwards (PupilSampling.jl:9, :27-30)
    xv, yv, k, f = orc.trace3d(S, 60.0, 0.0, 0.0, 0.0)
    assert f & orc.F_MISS and np.isnan(yv).all()
    assert np.allclose(k, [0.0, 0.0, 1.0])
    # plane-only system: straight line
    P = np.array([[np.inf, 0.0, 1.0], [np.inf, 10.0, 1.0], [np.inf, 0.0, 1.0]])
    xv, yv, k, f = orc.trace3d(P, 1.0, 2.0, 0.1, -0.2)
    assert f == 0 and np.allclose(yv, [1.0, 2.0]) and np.allclose(xv, [2.0, 0.0])
    # empty grid
    g = orc.grid_trace(np.vstack([S, [np.inf, 0, 1]]), np.zeros(0), np.zeros(3), 5, 10.3, 0.0)
    assert g["n_kept"] == 0 and g["ex"].size == 0


def test_seidel_aberrations(system, orc):
    """test/runtests.jl:160-229: per-surface Seidel contributions and totals vs the book table (0.25 wave)"""
    SE = COOKE["seidel"]
    d, per = orc.seidel(S, A, COOKE["h_prime"], dn=SE["dn"])
    assert d["f"] == system.f and d["EBFD"] == system.EBFD and d["stop"] == 5 and d["H"] == system.H
    alpha = 2 * system.marginal.u[-1] / KAT["constants"]["lambda_mm"]
    tol = SE["wave_scale_atol"]
    third = np.array(SE["third_order"])
    for j, div in enumerate(SE["third_order_divisors"]):            # spherical, coma, astigmatism, petzval, distortion
        assert np.max(np.abs(per[j] - alpha * third[:, j] / div)) < tol
    assert np.max(np.abs(per[5] - alpha * np.array(SE["PAC"]) / SE["PAC_divisor"])) < tol
    assert np.max(np.abs(per[6] - alpha * np.array(SE["PLC"]) / SE["PLC_divisor"])) < tol
    for name, div in SE["total_divisors"].items():
        assert abs(d[name] - alpha * SE[name] / div) < tol, name
    # Petzval curvature identity (test/runtests.jl:223-228)
    N = system.N
    rho = COOKE["h_prime"] ** 2 / (2 * (-8 * N ** 2 * d["W220P"] * KAT["constants"]["lambda_mm"]))
    n = S[:, 2]
    ptzc = -np.sum(system.phi / (n[1:] * n[:-1]))
    assert math.isclose(1 / rho, ptzc, rel_tol=1e-9)
