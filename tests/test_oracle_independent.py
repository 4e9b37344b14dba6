"""Pins the oracle's skew-ray arithmetic (orc_trace3d: src/PupilSampling.jl:1-65) off the meridional plane against an
INDEPENDENT 50-digit tracer written from vector geometry (oracle/independent_tracer.py: ray-conicoid quadratic in direction
cosines, gradient normal, vector Snell / mirror law -- none of the reference's sag / tilt / refract! formulas).  The
reference's own tests constrain x != 0 rays only through a spot RMS at +-0.07 (test/runtests.jl:364-372)."""
import numpy as np
import pytest

FIXTURES = ("COOKE", "DOUBLE_GAUSS", "DOUBLE_GAUSS_CONIC", "REFLECTIVE", "PARABOLA")


def fixture(ort, name):
    P = ort.prescriptions
    if name == "COOKE":
        S, K, amax, umax, img = P.COOKE["surfaces"], None, 12.0, 0.25, 60.0
    elif name.startswith("DOUBLE_GAUSS"):
        S, amax, umax, img = P.DOUBLE_GAUSS["surfaces"], 20.0, 0.2, 50.0
        K = np.where(np.isfinite(S[:, 0]), -0.3, 0.0) if name.endswith("CONIC") else None
    elif name == "REFLECTIVE":
        S, K, amax, umax, img = P.REFLECTIVE["surfaces"], None, 10.0, 0.05, -30.0
    else:
        S, K, amax, umax, img = P.PARABOLA["surfaces"][:, :3], P.PARABOLA["surfaces"][:, 3], 25.0, 0.05, -30.0
    ext = np.vstack([S, [np.inf, 0.0, S[-1, 2]]])                    # a plane behind the last surface, as full_trace appends one
    ext[-2, 1] = img
    return ext, (None if K is None else np.append(K, 0.0)), amax, umax


def compare(orc, ort, name, n_rays, seed=0):
    from oracle import independent_tracer as it
    ext, K, amax, umax = fixture(ort, name)
    rng = np.random.default_rng(seed)
    y, x = rng.uniform(-amax, amax, n_rays), rng.uniform(-amax, amax, n_rays)
    u, v = rng.uniform(-umax, umax, n_rays), rng.uniform(-umax, umax, n_rays)
    xv, yv, k, fl = orc.trace3d_batch(ext, y, x, u, v, K=K)
    worst_p = worst_d = 0.0
    n_ok = n_skip = 0
    for i in range(n_rays):
        xs, ys, d, st = it.trace(ext, y[i], x[i], u[i], v[i], K=K)
        if st != "ok" or fl[i] != 0:
            assert st != "ok" or (fl[i] & 2), (name, i, st, fl[i])   # a geometric miss must be a miss on both sides
            n_skip += 1
            continue
        px, py = np.array([float(a) for a in xs]), np.array([float(a) for a in ys])
        sc = max(np.max(np.abs(px)), np.max(np.abs(py)), 1.0)
        worst_p = max(worst_p, np.max(np.abs(px - xv[:, i])) / sc, np.max(np.abs(py - yv[:, i])) / sc)
        dd = np.array([float(a) for a in d])
        dd = dd * np.sign(dd[2])                                     # the reference's k is a line direction with k3 > 0
        worst_d = max(worst_d, float(np.max(np.abs(dd - k[:, i]))))
        n_ok += 1
    return dict(rays=n_rays, compared=n_ok, skipped=n_skip, max_position_err_rel=worst_p, max_direction_err=worst_d)


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_skew_rays_match_independent_50_digit_tracer(orc, ort, name):
    r = compare(orc, ort, name, 2500)
    assert r["compared"] > 2000
    assert r["max_position_err_rel"] < 2e-14 and r["max_direction_err"] < 2e-14, r
