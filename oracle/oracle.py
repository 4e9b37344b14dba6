"""ctypes binding of the CPU ORACLE (oracle/ort_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("ORT_ORACLE_LIB", os.path.join(_HERE, "libort_oracle.so"))      # the sanitizer build sets it

F_MISS, F_TIR, F_DOMAIN, F_CLIP, F_VIGN = 1, 2, 4, 8, 16

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


def build(force=False):
    src = os.path.join(_HERE, "ort_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_hypot.restype = C.c_double
        L.orc_hypot.argtypes = [C.c_double, C.c_double]
        L.orc_lens.restype = C.c_int
        L.orc_paraxial_trace.restype = C.c_int
        L.orc_trace2d.restype = C.c_uint
        L.orc_trace3d.restype = C.c_uint
        L.orc_grid_trace.restype = C.c_int64
        L.orc_compact.restype = C.c_int64
        L.orc_sum.restype = C.c_double
        L.orc_sigma.restype = C.c_double
        L.orc_mirror_stats.restype = C.c_double
        L.orc_max_threads.restype = C.c_int
        L.orc_trace3d_ext.restype = C.c_uint
        L.orc_opl_start.restype = C.c_double
        L.orc_grid_trace_ext.restype = C.c_int64
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_poly(coef=None):
    """EXTENSION: aspheric polynomial terms in coefficient form, coef[row][k] multiplies y^k (row 0 = object space,
    ignored); None clears them.  Process-global state read by the 2-D and 3-D tracers."""
    if coef is None:
        lib().orc_set_poly(C.c_int(0), C.c_int(0), None)
        return
    c = _d(coef)
    assert c.ndim == 2
    rc = lib().orc_set_poly(C.c_int(c.shape[0]), C.c_int(c.shape[1]), c.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise ValueError("orc_set_poly: too many rows / coefficients")


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _pu8(a):
    return None if a is None else a.ctypes.data_as(_u8p)


def max_threads():
    return lib().orc_max_threads()


def hypot(x, y):
    return lib().orc_hypot(float(x), float(y))


# ---------------------------------------------------------------- paraxial
def lens(surfaces):
    """Lens(surfaces) -> (tau, phi, n)   [src/RayTracing.jl:38-53]"""
    S = np.asarray(surfaces, dtype=np.float64)
    R, t, n = _d(S[:, 0]), _d(S[:, 1]), _d(S[:, 2])
    rows = len(R)
    tau, phi = np.empty(rows), np.empty(rows)
    k = lib().orc_lens(C.c_int(rows), _p(R), _p(t), _p(n), _p(tau), _p(phi))
    return tau[:k].copy(), phi[:k].copy(), n.copy()


def paraxial_trace(tau, phi, y, w, a=None, clip=False):
    """raytrace(lens, y, w, a; clip) -> (rt (k+1)x2 [y nu], clip_index)  [RayTracing.jl:127-143]"""
    tau, phi = _d(tau), _d(phi)
    k = len(tau)
    a_ = None if a is None else _d(a)
    rt = np.empty((2, k + 1))
    ci = lib().orc_paraxial_trace(C.c_int(k), _p(tau), _p(phi), _p(a_), C.c_int(int(clip)),
                                  C.c_double(y), C.c_double(w), _p(rt))
    return rt.T.copy(), ci


def paraxial_batch(tau, phi, y0, w0, a=None, clip=False, threads=0):
    tau, phi, y0, w0 = _d(tau), _d(phi), _d(y0), _d(w0)
    a_ = None if a is None else _d(a)
    N = len(y0)
    y, w = np.empty(N), np.empty(N)
    ci = np.zeros(N, dtype=np.int32)
    lib().orc_paraxial_batch(C.c_int(len(tau)), _p(tau), _p(phi), _p(a_), C.c_int(int(clip)),
                             C.c_int64(N), _p(y0), _p(w0), _p(y), _p(w),
                             ci.ctypes.data_as(_i32p), C.c_int(threads))
    return y, w, ci


# ---------------------------------------------------------------- transfer matrix
def transfer_matrix(tau, phi):
    """TransferMatrix(lens) -> 2x2 ndarray  [src/TransferMatrix.jl:1-6]"""
    tau, phi = _d(tau), _d(phi)
    M = np.empty(4)
    lib().orc_transfer_matrix(C.c_int(len(tau)), _p(tau), _p(phi), _p(M))
    return M.reshape(2, 2).T.copy()  # stored column-major


def _colmajor(M):
    return _d(np.asarray(M, dtype=np.float64).T.reshape(-1))


def transfer(M, v, tau, taup):
    Mc, v = _colmajor(M), _d(v)
    out = np.empty(2)
    lib().orc_transfer(_p(Mc), C.c_double(tau), C.c_double(taup), _p(v), _p(out))
    return out


def reverse_transfer(M, v, taup, tau):
    Mc, v = _colmajor(M), _d(v)
    out = np.empty(2)
    lib().orc_reverse_transfer(_p(Mc), C.c_double(taup), C.c_double(tau), _p(v), _p(out))
    return out


def transfer_batch(M, tau, taup, v_in, reverse=False, threads=0):
    """v_in: (N,2) rows [y, nu] -> (N,2)"""
    Mc = _colmajor(M)
    v_in = _d(v_in)
    N = v_in.shape[0]
    out = np.empty_like(v_in)
    lib().orc_transfer_batch(_p(Mc), C.c_double(tau), C.c_double(taup), C.c_int(int(reverse)),
                             C.c_int64(N), _p(v_in), _p(out), C.c_int(threads))
    return out


# ---------------------------------------------------------------- real rays
def _cols(surfaces, K=None):
    S = np.asarray(surfaces, dtype=np.float64)
    R, t, n = _d(S[:, 0]), _d(S[:, 1]), _d(S[:, 2])
    if K is None and S.shape[1] > 3:
        K = S[:, 3]
    Kc = None if K is None else _d(K)
    return R, t, n, Kc


def trace2d(surfaces, y, U, K=None, aspheric=False):
    """raytrace(surfaces, y, U, RealRay) -> (rt rows x 2 [y U], ts, flags)  [RayTracing.jl:145-173]"""
    R, t, n, Kc = _cols(surfaces, K)
    rows = len(R)
    if aspheric and Kc is None:
        Kc = np.zeros(rows)
    rt = np.empty((2, rows))
    ts = np.empty(rows)
    f = lib().orc_trace2d(C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), C.c_int(int(aspheric)),
                          C.c_double(y), C.c_double(U), _p(rt), _p(ts))
    return rt.T.copy(), ts, int(f)


def trace2d_batch(surfaces, y0, U0, K=None, aspheric=False, threads=0):
    R, t, n, Kc = _cols(surfaces, K)
    rows = len(R)
    if aspheric and Kc is None:
        Kc = np.zeros(rows)
    y0, U0 = _d(y0), _d(U0)
    N = len(y0)
    yo, Uo, ts = np.empty((rows, N)), np.empty((rows, N)), np.empty((rows, N))
    fl = np.zeros(N, dtype=np.uint8)
    lib().orc_trace2d_batch(C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), C.c_int(int(aspheric)),
                            C.c_int64(N), _p(y0), _p(U0), _p(yo), _p(Uo), _p(ts), _pu8(fl),
                            C.c_int(threads))
    return yo, Uo, ts, fl


def trace3d(surfaces, y, x, u, v, K=None):
    """3-D skew trace with SLOPES u=tan(U), v=tan(V) -> (xv, yv, k, flags)  [PupilSampling.jl:34-65]"""
    R, t, n, Kc = _cols(surfaces, K)
    rows = len(R)
    xv, yv, k = np.empty(rows - 1), np.empty(rows - 1), np.empty(3)
    f = lib().orc_trace3d(C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), C.c_double(y),
                          C.c_double(x), C.c_double(u), C.c_double(v), _p(xv), _p(yv), _p(k))
    return xv, yv, k, int(f)


def trace3d_batch(surfaces, y0, x0, u0, v0, K=None, threads=0):
    R, t, n, Kc = _cols(surfaces, K)
    rows = len(R)
    y0, x0, u0, v0 = _d(y0), _d(x0), _d(u0), _d(v0)
    N = len(y0)
    xv, yv, k = np.empty((rows - 1, N)), np.empty((rows - 1, N)), np.empty((3, N))
    fl = np.zeros(N, dtype=np.uint8)
    lib().orc_trace3d_batch(C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), C.c_int64(N), _p(y0),
                            _p(x0), _p(u0), _p(v0), _p(xv), _p(yv), _p(k), _pu8(fl),
                            C.c_int(threads))
    return xv, yv, k, fl


def trace3d_ld_batch(surfaces, y0, x0, u0, v0, K=None):
    """Extended-precision (80-bit) evaluation of the same trace: the 'truth' for conditioning checks."""
    R, t, n, Kc = _cols(surfaces, K)
    rows = len(R)
    y0, x0, u0, v0 = _d(y0), _d(x0), _d(u0), _d(v0)
    N = len(y0)
    xv, yv, k = np.empty((rows - 1, N)), np.empty((rows - 1, N)), np.empty((3, N))
    lib().orc_trace3d_ld_batch(C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), C.c_int64(N), _p(y0),
                               _p(x0), _p(u0), _p(v0), _p(xv), _p(yv), _p(k))
    return xv, yv, k


def grid_trace(ext_surfaces, ys, xs, stop, a_stop, h_prime, u=0.0, v=0.0, mode=0, ybar=0.0,
               z0=1.0, K=None, want=("ex", "ey", "r", "theta", "mask", "flags"), threads=0):
    """The hot loop of full_trace (PupilSampling.jl:115-138) on the EXTENDED surfaces.
    Returns dict of full-grid arrays (ny*nx, y outer / x inner) + 'n_kept'."""
    R, t, n, Kc = _cols(ext_surfaces, K)
    rows = len(R)
    ys, xs = _d(ys), _d(xs)
    ny, nx = len(ys), len(xs)
    NN = ny * nx
    out = {}
    for name in ("ex", "ey", "r", "theta"):
        out[name] = np.empty(NN) if name in want else None
    out["mask"] = np.zeros(NN, dtype=np.uint8) if "mask" in want else None
    out["flags"] = np.zeros(NN, dtype=np.uint8) if "flags" in want else None
    kept = lib().orc_grid_trace(
        C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), C.c_int(mode), C.c_double(u), C.c_double(v),
        C.c_double(ybar), C.c_double(z0), C.c_double(h_prime), C.c_int(ny), _p(ys), C.c_int(nx),
        _p(xs), C.c_int(stop), C.c_double(a_stop), _p(out["ex"]), _p(out["ey"]), _p(out["r"]),
        _p(out["theta"]), _pu8(out["mask"]), _pu8(out["flags"]), C.c_int(threads))
    out["n_kept"] = int(kept)
    return out


def opl_start(mode, n0, y, x, u, v, z0=1.0):
    return lib().orc_opl_start(C.c_int(mode), C.c_double(n0), C.c_double(y), C.c_double(x), C.c_double(u),
                               C.c_double(v), C.c_double(z0))


def trace3d_ext(surfaces, y, x, u, v, K=None, a=None, opl0=0.0, xc=0.0, yc=0.0, rr=0.0, truth=False):
    """EXTENSION (no reference counterpart): 3-D trace with OPL accumulation, per-surface apertures and an
    optional reference sphere.  -> (xv, yv, k, opl, flags); truth=True -> 80-bit OPL only."""
    R, t, n, Kc = _cols(surfaces, K)
    rows = len(R)
    a_ = None if a is None else _d(a)
    opl = C.c_double()
    if truth:
        lib().orc_trace3d_ext_ld(C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), C.c_double(y), C.c_double(x),
                                 C.c_double(u), C.c_double(v), C.c_double(opl0), C.c_double(xc), C.c_double(yc),
                                 C.c_double(rr), C.byref(opl))
        return opl.value
    xv, yv, k = np.empty(rows - 1), np.empty(rows - 1), np.empty(3)
    f = lib().orc_trace3d_ext(C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), _p(a_), C.c_double(y), C.c_double(x),
                              C.c_double(u), C.c_double(v), C.c_double(opl0), C.c_double(xc), C.c_double(yc),
                              C.c_double(rr), _p(xv), _p(yv), _p(k), C.byref(opl))
    return xv, yv, k, opl.value, int(f)


def grid_trace_ext(ext_surfaces, ys, xs, stop, a_stop, h_prime, u=0.0, v=0.0, mode=0, ybar=0.0, z0=1.0, K=None,
                   a=None, xc=0.0, yc=0.0, rr=0.0, opl_ref=0.0, opd_scale=1.0, threads=0):
    """EXTENSION: grid sweep with OPD output and per-surface aperture clipping."""
    R, t, n, Kc = _cols(ext_surfaces, K)
    rows = len(R)
    ys, xs = _d(ys), _d(xs)
    a_ = None if a is None else _d(a)
    NN = len(ys) * len(xs)
    out = {"ex": np.empty(NN), "ey": np.empty(NN), "opd": np.empty(NN), "mask": np.zeros(NN, dtype=np.uint8),
           "flags": np.zeros(NN, dtype=np.uint8)}
    kept = lib().orc_grid_trace_ext(
        C.c_int(rows), _p(R), _p(t), _p(n), _p(Kc), _p(a_), C.c_int(mode), C.c_double(u), C.c_double(v),
        C.c_double(ybar), C.c_double(z0), C.c_double(h_prime), C.c_double(xc), C.c_double(yc), C.c_double(rr),
        C.c_double(opl_ref), C.c_double(opd_scale), C.c_int(len(ys)), _p(ys), C.c_int(len(xs)), _p(xs),
        C.c_int(stop), C.c_double(a_stop), _p(out["ex"]), _p(out["ey"]), _p(out["opd"]), _pu8(out["mask"]),
        _pu8(out["flags"]), C.c_int(threads))
    out["n_kept"] = int(kept)
    return out


SEIDEL_FIELDS = ("f", "EBFD", "stop", "H", "W040", "W131", "W222", "W220P", "W311", "W020", "W111", "W220", "W220M",
                 "W220T", "nu_marginal", "nu_chief")


def seidel(surfaces, a, h_prime, lam=587.5618e-6, dn=None):
    """solve + aberrations of one prescription -> (dict of the 16 scalars, per-surface (7, k) table)
    [src/RayTracing.jl:38-53,208-221,246-263; src/SeidelAberrations.jl:6-53]"""
    R, t, n, _ = _cols(surfaces)
    rows = len(R)
    a = _d(a)
    dn_ = None if dn is None else _d(dn)
    out, per = np.empty(16), np.empty((7, rows - 1))
    rc = lib().orc_seidel(C.c_int(rows), _p(R), _p(t), _p(n), _p(a), C.c_double(h_prime), C.c_double(lam), _p(dn_),
                          _p(out), _p(per))
    if rc != 0:
        raise ValueError(f"orc_seidel failed ({rc})")
    return dict(zip(SEIDEL_FIELDS, out)), per


def seidel_candidates(RtnK, a, h_prime, lam=587.5618e-6, dn=None, threads=0):
    RtnK = _d(RtnK)
    Cn, four, rows = RtnK.shape
    a = _d(a)
    dn_ = None if dn is None else _d(dn)
    out = np.empty((Cn, 16))
    lib().orc_seidel_candidates(C.c_int(rows), C.c_int64(Cn), _p(RtnK), _p(a), C.c_double(h_prime), C.c_double(lam),
                                _p(dn_), _p(out), C.c_int(threads))
    return out


def compact(mask, arr):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    arr = _d(arr)
    out = np.empty(int(mask.sum()))
    c = lib().orc_compact(C.c_int64(len(mask)), _pu8(mask), _p(arr), _p(out))
    assert c == len(out)
    return out


def pairwise_sum(a):
    a = _d(a)
    return lib().orc_sum(C.c_int64(len(a)), _p(a))


def sigma(ex, ey):
    ex, ey = _d(ex), _d(ey)
    return lib().orc_sigma(C.c_int64(len(ex)), _p(ex), _p(ey))


def mirror_stats(ex, ey, r, theta):
    """PupilSampling.jl:139-146 -> (ex2, ey2, rho2, th2, RMS)"""
    ex, ey, r, theta = _d(ex), _d(ey), _d(r), _d(theta)
    n = len(ex)
    ex2, ey2, rho2, th2 = (np.empty(2 * n) for _ in range(4))
    rms = lib().orc_mirror_stats(C.c_int64(n), _p(ex), _p(ey), _p(r), _p(theta), _p(ex2),
                                 _p(ey2), _p(rho2), _p(th2))
    return ex2, ey2, rho2, th2, rms


def wavegrad(e, nu, lam=587.5618e-6):
    e = _d(e)
    out = np.empty_like(e)
    lib().orc_wavegrad(C.c_int64(len(e)), _p(e), C.c_double(nu), C.c_double(lam), _p(out))
    return out


def candidates(RtnK, ys, xs, stop, a_stop, h_prime, u, v=0.0, threads=0):
    """RtnK: (C, 4, rows).  Returns (C, 4): n_kept, mean_x, mean_y, RMS."""
    RtnK = _d(RtnK)
    Cn, four, rows = RtnK.shape
    assert four == 4
    ys, xs = _d(ys), _d(xs)
    out = np.empty((Cn, 4))
    lib().orc_candidates(C.c_int(rows), C.c_int64(Cn), _p(RtnK), C.c_double(u), C.c_double(v),
                         C.c_double(h_prime), C.c_int(len(ys)), _p(ys), C.c_int(len(xs)), _p(xs),
                         C.c_int(stop), C.c_double(a_stop), _p(out), C.c_int(threads))
    return out
