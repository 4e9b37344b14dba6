"""TEST INFRASTRUCTURE: a backend with the same methods as ort_b200.Context, answered by the CPU
oracle, so the host logic (solve, ray aiming, full_trace assembly, stats merge) can be tested
without a GPU.  The product never selects this backend; only tests inject it."""
import numpy as np

from oracle import oracle as orc


class OracleBackend:
    def __init__(self):
        self.S = None
        self.K = None
        self.rows = 0
        self.a = None

    def set_layout(self, surfaces, K=None):
        S = np.asarray(surfaces, dtype=np.float64)
        self.S = S[:, :3].copy()
        self.K = None if K is None else np.asarray(K, dtype=np.float64).copy()
        self.rows = S.shape[0]
        orc.set_poly(None)                          # like ort_set_layout: a new layout clears the polynomial terms

    def set_polynomials(self, coef=None):
        orc.set_poly(coef)

    def set_apertures(self, a):
        self.a = None if a is None else np.asarray(a, dtype=np.float64).copy()

    def trace2d_batch(self, y0, U0, aspheric=False):
        return orc.trace2d_batch(self.S, y0, U0, K=self.K, aspheric=aspheric)

    def trace3d_rays(self, y0, x0, u0, v0, arith=0, opl=False):
        xv, yv, k, fl = orc.trace3d_batch(self.S, y0, x0, u0, v0, K=self.K)
        if not opl:
            return xv, yv, k, fl
        ol = np.array([orc.trace3d_ext(self.S, y0[i], x0[i], u0[i], v0[i], K=self.K)[3] for i in range(len(y0))])
        return xv, yv, k, fl, ol

    def paraxial_batch(self, tau, phi, y0, w0, a=None, clip=False, arith=0, table=False):
        y, w, ci = orc.paraxial_batch(tau, phi, y0, w0, a=a, clip=clip)
        if not table:
            return y, w, ci
        k, N = len(tau), len(y0)
        ya, wa = np.empty((k + 1, N)), np.empty((k + 1, N))
        for j in range(N):
            rt, _ = orc.paraxial_trace(tau, phi, float(y0[j]), float(w0[j]), a=a, clip=clip)
            ya[:, j], wa[:, j] = rt[:, 0], rt[:, 1]
        return y, w, ci, ya, wa

    def seidel_candidates(self, RtnK, a, h_prime, lam=587.5618e-6, dn=None, per_surface=False):
        out = orc.seidel_candidates(RtnK, a, h_prime, lam=lam, dn=dn)
        if not per_surface:
            return out
        per = np.stack([orc.seidel(np.asarray(c)[:3].T, a, h_prime, lam=lam, dn=dn)[1] for c in RtnK])
        return out, per

    def aim_candidates(self, RtnK, a, h_prime, H, aspheric=False):
        from oracle import prelude as pre
        out = np.full((len(RtnK), 24), np.nan)
        for c, blk in enumerate(np.asarray(RtnK, dtype=np.float64)):
            S, K = blk[:3].T.copy(), blk[3].copy()
            sysm = pre.solve(S, a, h_prime)
            p = pre.full_trace_inputs(sysm, H, 64, K=K, aspheric=aspheric)
            out[c, :16] = (p.y1, p.y2, p.y_EP, p.u, p.h_prime, p.focus, p.stop, p.a_stop, p.EP_t,
                           p.U / abs(H) if H else np.nan, sysm.f, 0.0, p.nu, p.U, 0.0, 0.0)
        return out

    def trace3d_candidates_aimed(self, RtnK, aim, ny, nx, arith=1):
        out = np.empty((len(RtnK), 4))
        for c, blk in enumerate(np.asarray(RtnK, dtype=np.float64)):
            y1, y2, y_EP, u, hp, focus, stop, a_stop = aim[c, :8]
            ext = np.concatenate([blk, np.array([[np.inf], [0.0], [1.0], [0.0]])], axis=1)
            ext[1, -2] = focus
            out[c] = orc.candidates(ext[None], np.linspace(y1, y2, ny), np.linspace(0.0, y_EP, nx), int(stop), a_stop, hp, u)[0]
        return out

    def transfer_batch(self, M, tau, taup, v_in, reverse=False):
        return orc.transfer_batch(M, tau, taup, v_in, reverse=reverse)

    def trace3d_grid(self, fields, ys, xs, stop, a_stop, arith=1, compact=False,
                     want=("ex", "ey", "mask", "stats"), wavegrad=None, out=None, ext=0, opd_scale=1.0):
        from ort_b200 import STATS_DTYPE
        ys = np.asarray(ys, dtype=np.float64)
        nf, NN = len(fields), ys.shape[-1] * len(xs)
        res = {k: np.full((nf, NN), np.nan) for k in ("ex", "ey", "r", "theta", "opd")}
        res["mask"] = np.zeros((nf, NN), dtype=np.uint8)
        res["flags"] = np.zeros((nf, NN), dtype=np.uint8)
        stats = np.zeros(nf, dtype=STATS_DTYPE)
        for f, fld in enumerate(fields):
            g = orc.grid_trace(self.S, ys[f] if ys.ndim == 2 else ys, xs, stop, a_stop, fld.get("h_prime", 0.0), u=fld.get("u", 0.0),
                               v=fld.get("v", 0.0), mode=fld.get("mode", 0), ybar=fld.get("ybar", 0.0),
                               z0=fld.get("z0", 1.0), K=self.K)
            if ext or "opd" in want:
                a = None
                if ext & 2 and self.a is not None:
                    a = np.full(self.rows - 1, np.inf); a[:len(self.a)] = self.a
                ge = orc.grid_trace_ext(self.S, ys[f] if ys.ndim == 2 else ys, xs, stop, a_stop, fld.get("h_prime", 0.0),
                                        u=fld.get("u", 0.0), v=fld.get("v", 0.0), mode=fld.get("mode", 0),
                                        ybar=fld.get("ybar", 0.0), z0=fld.get("z0", 1.0), K=self.K, a=a,
                                        xc=fld.get("opd_xc", 0.0), yc=fld.get("opd_yc", 0.0), rr=fld.get("opd_radius", 0.0),
                                        opl_ref=fld.get("opl_ref", 0.0), opd_scale=opd_scale)
                g["mask"], g["flags"], g["opd"] = ge["mask"], ge["flags"], ge["opd"]
            else:
                g["opd"] = np.zeros(NN)
            m = g["mask"].astype(bool)
            n = int(m.sum())
            for k in ("ex", "ey", "r", "theta", "opd"):
                if compact:
                    res[k][f, :n] = g[k][m]
                else:
                    res[k][f] = g[k]
            res["mask"][f], res["flags"][f] = g["mask"], g["flags"]
            st = stats[f]
            st["n_kept"] = n
            if n:
                ex, ey = g["ex"][m], g["ey"][m]
                st["mean_x"], st["mean_y"] = ex.mean(), ey.mean()
                st["m2_x"], st["m2_y"] = ((ex - ex.mean()) ** 2).sum(), ((ey - ey.mean()) ** 2).sum()
                st["r_max"] = g["r"][m].max()
                st["mean_opd"] = g["opd"][m].mean()
                st["m2_opd"] = ((g["opd"][m] - g["opd"][m].mean()) ** 2).sum()
            for k, bit in (("n_miss", 1), ("n_tir", 2), ("n_domain", 4), ("n_clip", 8), ("n_vig", 16)):
                st[k] = int(((g["flags"] & bit) != 0).sum())
        res["stats"] = stats
        return res
