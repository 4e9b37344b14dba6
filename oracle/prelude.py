"""ORACLE prelude (TEST INFRASTRUCTURE ONLY): restatement of the reference's sequential host
prelude -- first-order solve and real-ray aiming -- needed to build full_trace inputs and to pin
the oracle against the reference's known-answer tests.  Scalar Python over the C oracle.

Follows (paths relative to /root/reference):
  src/Types.jl:37-50            ParaxialRay derived fields (u, z)
  src/RayTracing.jl:202-221     extend / trace_marginal_ray(lens, a)
  src/RayTracing.jl:246-263     trace_chief_ray(lens, stop, marginal, h')
  src/RayTracing.jl:302-323     _solve
  src/RayTracing.jl:117-125, 223-240, 265-296   real marginal / chief ray aiming
  src/PupilSampling.jl:67-83    trace_edge_rays (Optim.BFGS on |dy_stop|; restated as a secant
                                root find to the same root -- "parity unpinned", Optim.jl is not
                                under /root/reference)
  src/PupilSampling.jl:85-147   full_trace
  src/SeidelAberrations.jl:116-135  TSA
"""
import math
from types import SimpleNamespace as NS

import numpy as np

from . import oracle as orc

EPS = math.sqrt(np.finfo(np.float64).eps)  # const eps = sqrt(eps())  RayTracing.jl:1
LAMBDA = 587.5618e-6                       # SeidelAberrations.jl:2


def paraxial_ray(ynu, tau, n, fundamental):
    """ParaxialRay{T}(ynu, tau, n) -- Types.jl:37-50"""
    ynu = np.array(ynu, dtype=np.float64)
    y, nu = ynu[:, 0].copy(), ynu[:, 1].copy()
    n_ext = np.append(n, n[-1])
    m = min(len(nu), len(n_ext))
    u = nu[:m] / n_ext[:m]
    mt = min(len(tau), len(n_ext) - 2)
    t = np.asarray(tau[:mt]) * n_ext[:mt]
    if fundamental:
        t = np.append(t, -y[-2] / u[-2])
    z = np.cumsum(t)
    z0 = (z.min() - z.max()) * 0.1 if u[0] == 0.0 else -y[1] / u[0]
    z = np.concatenate([[z0], z])
    return NS(y=y, n=n_ext, u=u, yu=np.column_stack([y, u]), nu=nu, ynu=ynu, z=z)


def solve(surfaces, a, h_prime=-0.5, K=None):
    """solve(surfaces, a, h') -- RayTracing.jl:302-335.  `K` (optional) marks a Layout{Aspheric}."""
    S = np.array(surfaces, dtype=np.float64)[:, :3]
    a = np.asarray(a, dtype=np.float64)
    tau, phi, n = orc.lens(S)
    k = len(tau)
    # trace_marginal_ray(lens, a) :208-221
    rt, _ = orc.paraxial_trace(tau, phi, 1.0, 0.0)
    y, w = rt[:, 0], rt[:, 1]
    f = -1.0 / w[-1]
    EBFD = y[-1] * f
    sv = a / y[1:]
    stop = int(np.argmin(sv)) + 1           # findmin -> first minimal index, 1-based
    s = sv[stop - 1]
    mr = rt * s
    wf = mr[-1, 1]
    yf = mr[-1, 0] if wf == 0.0 else 0.0    # extend :202-206
    mr = np.vstack([mr, [yf, wf]])
    marginal = paraxial_ray(mr, tau, n, True)
    # trace_chief_ray(lens, stop, marginal, h') :246-263
    ys_ = marginal.y[1:-1]
    ynu_s = marginal.ynu[1:-1, :]
    y_stop = ys_[stop - 1]
    rt2, _ = orc.paraxial_trace(tau, phi, 0.0, 1.0)
    ynu2 = rt2[1:, :]
    y2_stop = ynu2[stop - 1, 0]
    nub = -marginal.nu[-1] * h_prime / ys_[0]
    cr = np.empty_like(marginal.ynu)
    cr[1:-1, :] = nub * (ynu2 - ynu_s * y2_stop / y_stop)
    cr[0, :] = (0.0, nub)
    cr[-1, :] = (h_prime, cr[-2, 1])
    chief = paraxial_ray(cr, tau, n, True)
    # _solve :302-323
    yb = chief.y[1]
    nub0 = chief.nu[0]
    nup = chief.nu[-1]
    ym = marginal.y[0]
    ypb = chief.y[-2]
    dp = EBFD - f
    d = (h_prime - nup * f - yb) / nub0
    EFFD = d - f
    PN = (n[-1] - n[0]) * f
    EP = NS(D=abs(ym) * 2, t=-yb / nub0)
    H = nub0 * ym
    XP = NS(D=abs(2 * H / nup), t=-ypb / nup)
    N = abs(f / EP.D)
    FOV = 2 * math.degrees(math.atan(abs(chief.u[0])))
    M = orc.transfer_matrix(tau, phi)
    return NS(f=f, EBFD=EBFD, EFFD=EFFD, N=N, FOV=FOV, stop=stop, EP=EP, XP=XP, marginal=marginal,
              chief=chief, H=H, P1=d, P2=dp, PN=PN, M=M, tau=tau, phi=phi, n=n, a=a,
              surfaces=S, K=None if K is None else np.asarray(K, dtype=np.float64),
              is_system=True)


def real_ray(surfaces, y, U, K=None, aspheric=False):
    """raytrace(surfaces, y, U, RealRay) -> RealRay{Tangential}; z = cumsum(ts) (Types.jl:61-63)"""
    rt, ts, fl = orc.trace2d(surfaces, y, U, K=K, aspheric=aspheric)
    return NS(y=rt[:, 0].copy(), u=rt[:, 1].copy(), yu=rt, n=np.asarray(surfaces)[:, 2].copy(),
              z=np.cumsum(ts), flags=fl)


def trace_marginal_ray_real(surfaces, system, K=None, aspheric=False, atol=EPS):
    """RayTracing.jl:223-240"""
    stop = system.stop
    y = system.marginal.y[0]
    u = 0.0
    a_stop = system.a[stop - 1]

    def loss(yy):
        r = real_ray(surfaces, yy, u, K, aspheric)
        return r, r.y[stop] - a_stop

    ray, d = loss(y)
    it = 0
    while abs(d) > atol:
        dy = loss(y + EPS)[1]
        y -= d * EPS / (dy - d)
        ray, d = loss(y)
        it += 1
        if it > 100:
            raise RuntimeError("marginal ray aiming did not converge")
    z = ray.z.copy()
    z[-1] = z[-2] - ray.y[-1] / math.tan(ray.u[-1])
    z = np.concatenate([[(z.min() - z.max()) * 0.1], z])
    yv = np.append(ray.y, 0.0)
    uv = np.append(ray.u, ray.u[-1])
    yu = np.vstack([ray.yu, [0.0, ray.u[-1]]])
    return NS(y=yv, u=uv, yu=yu, n=ray.n, z=z)


def trace_chief_ray_real(surfaces, system, K=None, is_layout=False, atol=EPS):
    """RayTracing.jl:265-296.  is_layout mirrors `surfaces isa Layout` (:272-277): the reversed
    system is then a Layout{Aspheric} (atan branch) whose K is reverse(K) -- shifted by one row
    against rev_R, exactly as the reference does it."""
    S = np.asarray(surfaces, dtype=np.float64)
    rows = S.shape[0]
    rev_R = -np.concatenate([[np.inf], S[:0:-1, 0]])
    rev_t = S[::-1, 1].copy()
    rev_n = S[::-1, 2].copy()
    m = system.marginal
    rev_t[0] = m.z[-1] - m.z[-2]
    rev = np.column_stack([rev_R, rev_t, rev_n])
    if is_layout:
        Kr = (np.zeros(rows) if K is None else np.asarray(K, dtype=np.float64))[::-1].copy()
        asph = True
    else:
        Kr, asph = None, False
    stop = rows - system.stop
    ybp = system.chief.y[-1]
    ubp = -system.chief.u[-1]

    def loss(uu):
        r = real_ray(rev, ybp, uu, Kr, asph)
        return r, r.y[stop]

    ray, ys = loss(ubp)
    it = 0
    while abs(ys) > atol:
        dy = loss(ubp + EPS)[1]
        ubp -= ys * EPS / (dy - ys)
        ray, ys = loss(ubp)
        it += 1
        if it > 100:
            raise RuntimeError("chief ray aiming did not converge")
    yb = np.concatenate([[0.0], ray.y[::-1]])
    yb[-1] = ybp
    ub = np.concatenate([-ray.u[::-1], [-ray.u[0]]])
    z = ray.z[-1] - ray.z[::-1]
    z[0] = -yb[1] / math.tan(ub[0]) + z[1]
    z = np.append(z, z[-1] - yb[-2] / math.tan(ub[-2]))
    return NS(y=yb, u=ub, yu=np.column_stack([yb, ub]), n=S[:, 2].copy(), z=z)


def _root(fun, x0, scale):
    """Secant/Newton root of fun near x0 (stands in for Optim.BFGS on abs(fun))."""
    x = x0
    fx = fun(x)
    for _ in range(60):
        if not math.isfinite(fx):
            raise RuntimeError("edge-ray aiming left the domain")
        if abs(fx) <= 4e-16 * scale:
            break
        h = EPS * max(1.0, abs(x))
        d = (fun(x + h) - fx) / h
        xn = x - fx / d
        fn = fun(xn)
        if abs(fn) >= abs(fx) and abs(fx) <= 1e-13 * scale:
            break
        x, fx = xn, fn
    return x


def trace_edge_rays(surfaces, y1, y2, U, stop, a_stop, K=None, aspheric=False):
    """PupilSampling.jl:67-83"""
    def ys(y):
        return real_ray(surfaces, y, U, K, aspheric).y[stop]
    r1 = _root(lambda y: ys(y) - a_stop, y1, a_stop)
    r2 = _root(lambda y: ys(y) + a_stop, y2, a_stop)
    return r1, r2


def _two_sum_by_magnitude(x, y):
    """Base.add12"""
    if abs(y) > abs(x):
        x, y = y, x
    h = x + y
    return h, (x - h) + y


def jl_range(start, stop, length):
    """collect(range(start, stop, length)) in Float64 -- Julia Base (base/twiceprecision.jl: range_start_stop_length,
    _linspace, unsafe_getindex of a TwicePrecision StepRangeLen; Julia 1.10).  That source is NOT under /root/reference
    and Julia cannot run here: restated from its published algorithm, scalar loop, and pinned only by the properties it
    is designed for (end points exact, every element within 1 ulp of -- almost always equal to -- the correctly rounded
    exact interpolation; tests/test_host_logic.py).  End points that are exact small rationals take Julia's integer
    branch, whose result is the correctly rounded rational: evaluated here with fractions.Fraction."""
    import struct
    from fractions import Fraction
    n, a, b = int(length), float(start), float(stop)
    if n == 1:
        return np.array([a])
    if a == b:
        return np.full(n, a)

    def rat_ok(x):                      # Base.rat finds x exactly with numerator, denominator <= maxintfloat(Float32)
        y, p0, q0, p1, q1 = x, 1, 0, 0, 1
        while abs(y) <= 16777216.0:
            f = int(y)
            y -= f
            p0, p1 = f * p0 + p1, p0
            q0, q1 = f * q0 + q1, q0
            if max(abs(p0), abs(q0)) > 16777216:
                return False
            if q0 != 0 and p0 / q0 == x:
                return True
            if y == 0.0:
                return False
            y = 1.0 / y
        return False
    if rat_ok(a) and rat_ok(b):
        fa, fb = Fraction(a), Fraction(b)
        return np.array([float(fa + (fb - fa) * Fraction(i, n - 1)) for i in range(n)])
    d = b - a
    tmin = -(a / d)
    imin = int(np.rint(tmin * (n - 1) + 1))
    if 1 < imin < n:
        t = (imin - 1) / (n - 1)
        ref = (1 - t) * a + t * b
        step = (ref - a) / (imin - 1) if imin - 1 < n - imin else (b - ref) / (n - imin)
    elif imin <= 1:
        imin, ref, step = 1, a, d / (n - 1)
    else:
        imin, ref, step = n, b, d / (n - 1)
    nb = min(27, int(math.ceil(math.log2(max(imin - 1, n - imin)))) + 1)
    bits = struct.unpack("<Q", struct.pack("<d", step))[0] & (0xFFFFFFFFFFFFFFFF << nb) & 0xFFFFFFFFFFFFFFFF
    step_hi = struct.unpack("<d", struct.pack("<Q", bits))[0]
    x1h, x1l = _two_sum_by_magnitude((1 - imin) * step_hi, ref)
    x2h, x2l = _two_sum_by_magnitude((n - imin) * step_hi, ref)
    ea, eb = (a - x1h) - x1l, (b - x2h) - x2l
    step_lo = (eb - ea) / (n - 1)
    ref_lo = ea - (1 - imin) * step_lo
    out = np.empty(n)
    for i in range(1, n + 1):
        u = i - imin
        xh, xl = _two_sum_by_magnitude(ref, u * step_hi)
        out[i - 1] = xh + (xl + (u * step_lo + ref_lo))
    return out


def full_trace_inputs(system, H, k_rays=64, focus=None, K=None, aspheric=False):
    """Host prelude of full_trace (PupilSampling.jl:85-122) for a System: everything the hot
    loop needs.  `aspheric` mirrors Layout{Aspheric} dispatch of the 2-D tracer."""
    H = abs(H)
    if not H <= 1.0:
        raise ValueError("Domain: |H| <= 1.0")
    S = system.surfaces
    if focus is None:
        focus = system.marginal.z[-1] - system.marginal.z[-2]
    stop = system.stop
    a_stop = abs(system.a[stop - 1])
    chief = trace_chief_ray_real(S, system, K=K, is_layout=True)      # system.layout isa Layout
    marg = trace_marginal_ray_real(S, system, K=K, aspheric=aspheric)
    EP_t = chief.z[0]
    Ubar = chief.u[0]
    U = H * Ubar
    u = math.tan(U)
    y_EP = abs(marg.y[0])
    y1, y2 = y_EP - u * EP_t, -y_EP - u * EP_t
    y1, y2 = trace_edge_rays(S, y1, y2, U, stop, a_stop, K=K, aspheric=aspheric)
    h_prime = u * system.f
    ext = np.vstack([S[:, :3], [np.inf, 0.0, 1.0]])
    Kx = np.append(np.zeros(S.shape[0]) if K is None else np.asarray(K, dtype=np.float64), 0.0)
    ext[-2, 1] = focus
    ys = jl_range(y1, y2, k_rays)                       # :121
    xs = jl_range(0.0, y_EP, k_rays // 2)               # :122
    return NS(ext=ext, K=Kx, ys=ys, xs=xs, u=u, v=math.tan(0.0), U=U, h_prime=h_prime, stop=stop,
              a_stop=a_stop, focus=focus, y1=y1, y2=y2, y_EP=y_EP, EP_t=EP_t,
              nu=system.marginal.nu[-1], H=H)


def full_trace(system, H, k_rays=64, focus=None, K=None, aspheric=False, threads=0):
    """full_trace(system, H, k_rays, focus) -> RealRayError-like namespace (PupilSampling.jl:85-147)"""
    p = full_trace_inputs(system, H, k_rays, focus, K, aspheric)
    g = orc.grid_trace(p.ext, p.ys, p.xs, p.stop, p.a_stop, p.h_prime, u=p.u, v=p.v, K=p.K,
                       threads=threads)
    m = g["mask"]
    ex, ey = orc.compact(m, g["ex"]), orc.compact(m, g["ey"])
    r, th = orc.compact(m, g["r"]), orc.compact(m, g["theta"])
    ex2, ey2, rho2, th2, rms = orc.mirror_stats(ex, ey, r, th)
    return NS(x=ex2, y=ey2, nu=p.nu, r=rho2, t=th2, H=p.H, RMS=rms, inputs=p, grid=g)


def tsa(surfaces, system, k_rays=22):
    """TSA(surfaces, system, k_rays) -- SeidelAberrations.jl:116-135 (plain-matrix surfaces)."""
    pm = system.marginal
    rm = trace_marginal_ray_real(surfaces, system)
    rc = trace_chief_ray_real(surfaces, system)
    XP_t = rc.z[-1] - rc.z[-2]
    y_EP = jl_range(rm.y[0] / k_rays, rm.y[0], k_rays)
    y_XP = np.empty(k_rays)
    eps_ = np.empty(k_rays)
    BFD = pm.z[-1] - pm.z[-2]
    t = BFD - (rm.z[-2] - pm.z[-2])                       # surface_to_focus :105, sag :93-95
    y_XP[-1] = rm.y[-2] + math.tan(rm.u[-1]) * XP_t
    eps_[-1] = rm.y[-2] + math.tan(rm.u[-2]) * t          # transfer(ray::RealRay{<:Fundamental}) :107
    for i in range(k_rays - 1):
        ray = real_ray(surfaces, y_EP[i], 0.0)
        t = BFD - (ray.z[-2] - ray.z[-1])                 # sag(ray::RealRay{Tangential}) :91
        y_XP[i] = ray.y[-1] + math.tan(ray.u[-1]) * XP_t
        eps_[i] = ray.y[-1] + math.tan(ray.u[-1]) * t
    return y_XP, eps_


def _isapprox(x, y):
    """Julia isapprox for Float64: rtol = sqrt(eps), atol = 0, elementwise"""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return (x == y) | (np.isfinite(x) & np.isfinite(y) & (np.abs(x - y) <= EPS * np.maximum(np.abs(x), np.abs(y))))


def _jl_minimum(v):
    v = np.asarray(v, dtype=np.float64)
    return np.nan if np.any(np.isnan(v)) else float(np.min(v))


def vignetting(system, a=None):
    """vignetting(system, a) -- src/Vignetting.jl:1-30.  Index lists are 1-based like the reference's findall."""
    a = np.asarray(system.a if a is None else a, dtype=np.float64)
    stop = system.stop
    yb = np.abs(system.chief.y[1:-1])                     # :3
    y = np.abs(system.marginal.y[1:-1])                   # :4
    k = len(a)
    M = np.empty((k, 5))
    M[:, 0] = a
    M[:, 1] = y
    M[:, 2] = y + yb
    M[:, 3] = yb
    M[:, 4] = yb - y
    with np.errstate(invalid="ignore", divide="ignore"):
        M[M[:, 3] < y, 3] = np.nan                        # :12
        M[M[:, 4] < y, 4] = np.nan                        # :13
        a_unvig = (a >= M[:, 2]) | _isapprox(a, M[:, 2])  # :14
        un = bool(np.all(a_unvig))
        min_un = _jl_minimum([(a[i] - y[i]) / yb[i] for i in range(k) if i != stop - 1])    # :17
        min_half = _jl_minimum(a / yb)                    # :18
        min_full = _jl_minimum((a + y) / yb)              # :19
        FOV = np.empty((3, 3))
        for i, sc in enumerate((min_un, min_half, min_full)):
            ub = abs(system.chief.u[0] * sc)
            FOV[i] = (2 * (math.atan(ub) * (180.0 / math.pi)), ub, abs(system.chief.y[-1] * sc))   # :22-25
        limit = np.nonzero((a < M[:, 1]) & ~_isapprox(a, M[:, 2]))[0] + 1                  # :27
        full = np.nonzero(a <= M[:, 4])[0] + 1                                             # :28
        partial = np.array([i for i in np.nonzero(~a_unvig)[0] + 1 if i not in set(full)], dtype=np.int64)   # :29
    return NS(M=M, FOV=FOV, un=un, limit=limit, partial=partial, full=full)
