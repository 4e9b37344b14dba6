"""EXTENSION rows of SURVEY.md section 8 (f3 per-surface aperture clipping, f4 OPL/OPD accumulation): no
reference counterpart, so they are pinned by physics -- a parabolic mirror images an axial point
perfectly, the rho^4 coefficient of the on-axis OPD map must reproduce the Seidel W040 the reference's
tests take from the book, OPL is rotationally symmetric on axis -- and by the 80-bit evaluation."""
import math

import numpy as np
import pytest

from oracle_backend import OracleBackend

LAM = 587.5618e-6


@pytest.fixture(scope="module")
def be(orc):
    return OracleBackend()


def test_parabola_has_zero_opd(ort, be):
    P = ort.prescriptions.PARABOLA
    s = ort.solve(ort.Layout(P["surfaces"]), P["a"], P["h"], backend=be)
    w = ort.wavefront(s.layout, s, [0.0], 64, backend=be)[0]
    assert w.rms < 1e-9 and w.pv < 1e-9                          # waves; OPL ~ 10 mm = 1.7e4 waves


def test_on_axis_w040_matches_seidel(ort, be):
    """test/runtests.jl:184,215: W040 = alpha * (-0.186575) / 8 waves, alpha = 2 u' / lambda"""
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=be)
    w = ort.wavefront(s.layout, s, [0.0], 64, backend=be)[0]
    e = ort.full_trace(s, 0.0, backend=be)
    n = len(w.opd) // 2
    assert len(e.x) == len(w.opd) and np.array_equal(e.x, w.x)  # same kept rays, same order
    rho = e.r[:n]
    A = np.column_stack([rho ** 2, rho ** 4, rho ** 6, rho ** 8])
    c = np.linalg.lstsq(A, w.opd[:n], rcond=None)[0]
    w040_seidel = 2 * s.marginal.u[-1] / LAM * (-0.186575) / 8
    assert abs(c[1] / w040_seidel - 1) < 0.03                   # third-order theory vs real rays
    assert abs(c[0]) < 0.1                                      # paraxial focus: no defocus term


def test_opl_truth_and_symmetry(ort, orc, pre):
    P = ort.prescriptions.DOUBLE_GAUSS
    so = pre.solve(P["surfaces"], P["a"], P["h"])
    p = pre.full_trace_inputs(so, 0.0, 16)
    rng = np.random.default_rng(4)
    for _ in range(50):
        r, th = rng.uniform(0, 0.9 * p.y_EP), rng.uniform(0, 2 * math.pi)
        y, x = r * math.cos(th), r * math.sin(th)
        _, _, _, opl, f = orc.trace3d_ext(p.ext, y, x, 0.0, 0.0, K=p.K, rr=60.0)
        truth = orc.trace3d_ext(p.ext, y, x, 0.0, 0.0, K=p.K, rr=60.0, truth=True)
        _, _, _, opl_m, _ = orc.trace3d_ext(p.ext, r, 0.0, 0.0, 0.0, K=p.K, rr=60.0)
        assert f == 0 and abs(opl / truth - 1) < 1e-14
        assert abs(opl / opl_m - 1) < 1e-14                      # rotational symmetry on axis


def test_vignetting_by_surface_apertures(ort, be):
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=be)
    w0, w1 = ort.wavefront(s.layout, s, [0.0, 1.0], 64, backend=be)
    v0, v1 = ort.wavefront(s.layout, s, [0.0, 1.0], 64, backend=be, vignette=True)
    assert len(v0.opd) == len(w0.opd)      # on axis the stop limits: only rays already outside the stop are vignetted
    assert int(v1.stats["n_vig"][0]) > 0 and len(v1.opd) < len(w1.opd)       # at full field the rims vignette
    # the surviving rays carry identical values: every (ex, ey, opd) triple of the vignetted run occurs unvignetted
    n1, nv = len(w1.opd) // 2, len(v1.opd) // 2
    full = {(a, b): c for a, b, c in zip(w1.x[:n1], w1.y[:n1], w1.opd[:n1])}
    assert all(full.get((a, b)) == c for a, b, c in zip(v1.x[:nv], v1.y[:nv], v1.opd[:nv]))
