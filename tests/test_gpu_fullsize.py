"""BASELINE-size checks (double-Gauss, 16 Mi rays per field) through size-independent properties:
FAST vs STRICT mask equality and 1e-12 agreement, statistics vs a torch recomputation from the
outputs, ordered compaction vs boolean indexing, row-sharded merge vs single sweep, idempotence,
and an oracle spot-check on a strided subsample of the same grid."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-12
NY, NX = 5792, 2896


@pytest.fixture(scope="module")
def dg(ctx, ort):
    P = ort.prescriptions.DOUBLE_GAUSS
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=ctx)
    p = ort.host._full_trace_setup(s.layout, s, [0.0, 1.0], 64, None, ctx)
    ctx.set_layout(p["ext"], p["K"])
    return s, p


def _sweep(ctx, ort, p, j, ys, xs, arith, compact=False, want=("ex", "ey", "mask", "flags", "stats")):
    dev = torch.device("cuda", 0)
    d_ys, d_xs = torch.from_numpy(ys).to(dev), torch.from_numpy(xs).to(dev)
    NN = len(ys) * len(xs)
    bufs = {k: torch.empty(NN, dtype=torch.float64, device=dev) for k in ("ex", "ey", "r", "theta") if k in want}
    for k in ("mask", "flags"):
        if k in want:
            bufs[k] = torch.empty(NN, dtype=torch.uint8, device=dev)
    stats = torch.zeros(ort.STATS_BYTES, dtype=torch.uint8, device=dev)
    ptrs = {k: v.data_ptr() for k, v in bufs.items()}
    ptrs["stats"] = stats.data_ptr()
    fld = dict(u=float(p["u"][j]), v=0.0, h_prime=float(p["h_prime"][j]))
    ctx.trace3d_grid_dev([fld], d_ys.data_ptr(), len(ys), d_xs.data_ptr(), len(xs), p["stop"], p["a_stop"], ptrs,
                         stream=torch.cuda.current_stream().cuda_stream, arith=arith, compact=compact)
    torch.cuda.synchronize()
    st = np.frombuffer(stats.cpu().numpy().tobytes(), dtype=ort.STATS_DTYPE)[0]
    return bufs, st


@pytest.mark.parametrize("j", [0, 1])
def test_fullsize_fast_vs_strict_and_stats(ctx, ort, orc, dg, j):
    s, p = dg
    ys = np.linspace(p["y1"][j], p["y2"][j], NY)
    xs = np.linspace(0.0, p["y_EP"], NX)
    bf, sf = _sweep(ctx, ort, p, j, ys, xs, ort.FAST)
    bs, ss = _sweep(ctx, ort, p, j, ys, xs, ort.STRICT)
    # clip mask and flags bit-exact between FAST and STRICT over all 16.8 M rays
    assert torch.equal(bf["mask"], bs["mask"]) and torch.equal(bf["flags"], bs["flags"])
    scale = max(abs(float(p["h_prime"][j])), p["y_EP"])
    m = bs["mask"].bool()
    assert float((bf["ex"][m] - bs["ex"][m]).abs().max()) / scale < TOL
    assert float((bf["ey"][m] - bs["ey"][m]).abs().max()) / scale < TOL
    assert torch.equal(torch.isnan(bf["ex"]), torch.isnan(bs["ex"]))
    # statistics == recomputation from the outputs (float64 torch reductions on the device)
    for b, st, tol in ((bs, ss, TOL), (bf, sf, TOL)):
        mm = b["mask"].bool()
        n = int(mm.sum())
        assert int(st["n_kept"]) == n
        ex, ey = b["ex"][mm], b["ey"][mm]
        assert abs(float(ex.mean()) - st["mean_x"]) < tol * scale and abs(float(ey.mean()) - st["mean_y"]) < tol * scale
        m2x = float(((ex - ex.mean()) ** 2).sum())
        m2y = float(((ey - ey.mean()) ** 2).sum())
        assert abs(m2x / st["m2_x"] - 1) < 1e-10 and abs(m2y / st["m2_y"] - 1) < 1e-10
        assert int(st["n_clip"]) == int(((b["flags"] & 8) != 0).sum())
    assert abs(ort.rms_from_stats(sf) - ort.rms_from_stats(ss)) < TOL * scale
    # oracle spot-check: every 97th row of the same grid, bit-exact against STRICT
    rows = np.arange(0, NY, 97)
    g = orc.grid_trace(p["ext"], ys[rows], xs, p["stop"], p["a_stop"], float(p["h_prime"][j]), u=float(p["u"][j]),
                       v=0.0, K=p["K"], want=("ex", "ey", "mask", "flags"))
    sel = (torch.from_numpy(rows).to(bs["ex"].device)[:, None] * NX + torch.arange(NX, device=bs["ex"].device)[None, :]).reshape(-1)
    assert np.array_equal(bs["mask"][sel].cpu().numpy(), g["mask"])
    gx, gy = bs["ex"][sel].cpu().numpy(), bs["ey"][sel].cpu().numpy()
    ok = ~np.isnan(g["ex"])
    assert np.array_equal(gx[ok].view(np.uint64), g["ex"][ok].view(np.uint64))
    assert np.array_equal(gy[ok].view(np.uint64), g["ey"][ok].view(np.uint64))
    assert np.abs(bf["ex"][sel].cpu().numpy()[ok] - g["ex"][ok]).max() / scale < TOL


def test_fullsize_compaction_and_sharding(ctx, ort, dg):
    s, p = dg
    j = 1
    ys = np.linspace(p["y1"][j], p["y2"][j], NY)
    xs = np.linspace(0.0, p["y_EP"], NX)
    full, st = _sweep(ctx, ort, p, j, ys, xs, ort.FAST, want=("ex", "ey", "r", "theta", "mask", "stats"))
    comp, stc = _sweep(ctx, ort, p, j, ys, xs, ort.FAST, compact=True, want=("ex", "ey", "r", "theta", "mask", "stats"))
    n = int(st["n_kept"])
    m = full["mask"].bool()
    assert int(stc["n_kept"]) == n == int(m.sum())
    for k in ("ex", "ey", "r", "theta"):                       # ordered compaction == boolean indexing
        assert torch.equal(comp[k][:n], full[k][m]), k
    assert abs(float(full["r"][m].max()) - st["r_max"]) <= 1e-12 * p["a_stop"]
    # idempotence: the same sweep twice is bit-identical (fixed-order reductions, no atomics)
    again, st2 = _sweep(ctx, ort, p, j, ys, xs, ort.FAST, want=("ex", "ey", "mask", "stats"))
    assert torch.equal(again["ex"], full["ex"]) and st2.tobytes() == st.tobytes()
    # row-sharded sweep (8 "ranks"), records merged in rank order == single sweep
    recs = np.zeros(8, dtype=ort.STATS_DTYPE)
    cat = []
    for r in range(8):
        lo, hi = ort.distributed.shard_rows(NY, r, 8)
        b, recs[r] = _sweep(ctx, ort, p, j, ys[lo:hi], xs, ort.FAST, want=("ex", "mask", "stats"))
        cat.append(b["ex"])
    mg = ort.merge_stats(recs)
    assert int(mg["n_kept"]) == n and mg["r_max"] == st["r_max"]
    assert torch.equal(torch.cat(cat), full["ex"])
    scale = max(abs(float(p["h_prime"][j])), p["y_EP"])
    assert abs(mg["mean_y"] - st["mean_y"]) < TOL * scale
    assert abs(ort.rms_from_stats(mg) / ort.rms_from_stats(st) - 1) < 1e-11


# ------------------------------------------------------------------------------------------------
# BASELINE configs 3, 4, 5 at their full sizes, through size-independent properties
# ------------------------------------------------------------------------------------------------
def test_config3_one_billion_rays_checksum_of_checksums(ctx, ort, dg):
    """config 3: 44722 x 22361 = 1.000e9 rays of one field on ONE GPU (stats + 1 GB mask): the mask sums to n_kept, and
    the sweep equals its 8 row shards merged in rank order (what 8 GPUs compute), counts exactly."""
    s, p = dg
    j = 1
    ny, nx = 44722, 22361
    ys = np.linspace(p["y1"][j], p["y2"][j], ny)
    xs = np.linspace(0.0, p["y_EP"], nx)
    dev = torch.device("cuda", 0)
    d_ys, d_xs = torch.from_numpy(ys).to(dev), torch.from_numpy(xs).to(dev)
    mask = torch.empty(ny * nx, dtype=torch.uint8, device=dev)
    stats = torch.zeros(ort.STATS_BYTES, dtype=torch.uint8, device=dev)
    fld = dict(u=float(p["u"][j]), v=0.0, h_prime=float(p["h_prime"][j]))
    st_ = torch.cuda.current_stream().cuda_stream

    def run(lo, hi, ptrs):
        ctx.trace3d_grid_dev([fld], d_ys.data_ptr() + 8 * lo, hi - lo, d_xs.data_ptr(), nx, p["stop"], p["a_stop"], ptrs,
                             stream=st_, arith=ort.FAST)
        torch.cuda.synchronize()
        return np.frombuffer(stats.cpu().numpy().tobytes(), dtype=ort.STATS_DTYPE)[0].copy()

    whole = run(0, ny, dict(mask=mask.data_ptr(), stats=stats.data_ptr()))
    assert int(mask.sum(dtype=torch.int64)) == int(whole["n_kept"]) > 0.7 * ny * nx
    assert int(whole["n_miss"]) == 0 and int(whole["n_tir"]) == 0
    recs = np.zeros(8, dtype=ort.STATS_DTYPE)
    for r in range(8):
        lo, hi = ort.distributed.shard_rows(ny, r, 8)
        recs[r] = run(lo, hi, dict(stats=stats.data_ptr()))
    mg = ort.merge_stats(recs)
    assert int(mg["n_kept"]) == int(whole["n_kept"]) and int(mg["n_clip"]) == int(whole["n_clip"])
    assert mg["r_max"] == whole["r_max"]
    assert abs(ort.rms_from_stats(mg) / ort.rms_from_stats(whole) - 1) < 1e-11
    # the billion-ray spot agrees with the 16 Mi-ray spot of the same field to sampling accuracy
    b, st16 = _sweep(ctx, ort, p, j, np.linspace(p["y1"][j], p["y2"][j], NY), np.linspace(0.0, p["y_EP"], NX), ort.FAST,
                     want=("stats",))
    assert abs(ort.rms_from_stats(whole) / ort.rms_from_stats(st16) - 1) < 2e-3


def test_config4_one_billion_paraxial_rays(ctx, orc, ort):
    """config 4: 40-row Lens, 1e9 rays device-resident.  Properties: linearity of the y-nu trace, the 2x2 transfer matrix
    applied to the same rays gives the same answer, transfer then reverse_transfer is the identity; and an oracle
    spot-check (bit-exact in STRICT) on a strided subsample."""
    N = 1_000_000_000
    L = ort.make_lens(ort.prescriptions.zoom20())
    tau, phi = L.tau, L.phi
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(42)
    y0 = (torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 20.0 - 10.0)
    w0 = (torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 0.4 - 0.2)
    y, w = torch.empty_like(y0), torch.empty_like(w0)
    for arith in (ort.STRICT, ort.FAST):
        ctx.paraxial_batch_dev(tau, phi, N, y0.data_ptr(), w0.data_ptr(), y.data_ptr(), w.data_ptr(), arith=arith)
        torch.cuda.synchronize()
        idx = torch.arange(0, N, 997_003, device=dev)
        ry, rw, _ = orc.paraxial_batch(tau, phi, y0[idx].cpu().numpy(), w0[idx].cpu().numpy())
        gy, gw = y[idx].cpu().numpy(), w[idx].cpu().numpy()
        if arith == ort.STRICT:
            assert np.array_equal(gy, ry) and np.array_equal(gw, rw)
        else:
            assert np.max(np.abs(gy - ry)) < 1e-12 * np.max(np.abs(ry)) and np.max(np.abs(gw - rw)) < 1e-12 * np.max(np.abs(rw))
    # linearity on the first 64 Mi rays: trace(2 a - 3 b) = 2 trace(a) - 3 trace(b)
    n = 1 << 26
    a_y, a_w, b_y, b_w = y0[:n], w0[:n], y0[n:2 * n], w0[n:2 * n]
    cy, cw = 2.0 * a_y - 3.0 * b_y, 2.0 * a_w - 3.0 * b_w
    oy, ow = torch.empty_like(cy), torch.empty_like(cw)
    ctx.paraxial_batch_dev(tau, phi, n, cy.data_ptr(), cw.data_ptr(), oy.data_ptr(), ow.data_ptr(), arith=ort.FAST)
    torch.cuda.synchronize()
    scale = float(y[:2 * n].abs().max())
    assert float((oy - (2.0 * y[:n] - 3.0 * y[n:2 * n])).abs().max()) < 1e-11 * scale
    del cy, cw, oy, ow
    # matrix == trace, and reverse(transfer) == identity, on all 1e9 rays
    M = ort.transfer_matrix(L)
    v = torch.stack([y0, w0], dim=1).contiguous()
    del y0, w0
    vo = torch.empty_like(v)
    ctx.transfer_batch_dev(M, 0.0, 0.0, N, v.data_ptr(), vo.data_ptr())
    torch.cuda.synchronize()
    sy, sw = float(y.abs().max()), float(w.abs().max())
    assert float((vo[:, 0] - y).abs().max()) < 1e-10 * sy and float((vo[:, 1] - w).abs().max()) < 1e-10 * sw
    del y, w
    back = torch.empty_like(v)
    ctx.transfer_batch_dev(M, 0.0, 0.0, N, vo.data_ptr(), back.data_ptr(), reverse=True)
    torch.cuda.synchronize()
    assert float((back - v).abs().max()) < 1e-9


def test_config5_candidate_population(ctx, orc, ort):
    """config 5: 65 536 triplet prescriptions x 4096 rays, shared grid and per-candidate aimed grid.  Properties: a
    candidate's result does not depend on its position in the batch (permutation), duplicates agree bit for bit, FAST and
    STRICT agree (counts exactly); and an oracle spot-check on 12 random candidates."""
    P = ort.prescriptions.COOKE
    C = 65536
    base = ort.prescriptions.perturbed_triplets(C)
    rng = np.random.default_rng(5)
    perm = rng.permutation(C)
    base[-1] = base[0]                                              # a duplicate
    aim = ctx.aim_candidates(base, P["a"], P["h"], 0.7)
    assert np.all(aim[:, 11] == 0.0) and aim[-1].tobytes() == aim[0].tobytes()
    spot = ctx.trace3d_candidates_aimed(base, aim, 64, 64, arith=ort.FAST)
    spot_p = ctx.trace3d_candidates_aimed(base[perm], aim[perm], 64, 64, arith=ort.FAST)
    assert spot_p.tobytes() == spot[perm].tobytes() and spot[-1].tobytes() == spot[0].tobytes()
    strict = ctx.trace3d_candidates_aimed(base, aim, 64, 64, arith=ort.STRICT)
    assert np.array_equal(strict[:, 0], spot[:, 0])
    assert np.max(np.abs(strict[:, 3] / spot[:, 3] - 1)) < 1e-10
    for c in rng.choice(C, 12, replace=False):
        y1, y2, y_EP, u, hp, focus, stop, a_stop = aim[c, :8]
        ext = np.concatenate([base[c], np.array([[np.inf], [0.0], [1.0], [0.0]])], axis=1)
        ext[1, -2] = focus
        ref = orc.candidates(ext[None], np.linspace(y1, y2, 64), np.linspace(0.0, y_EP, 64), int(stop), a_stop, hp, u)[0]
        assert strict[c, 0] == ref[0] and abs(strict[c, 3] / ref[3] - 1) < 1e-11
    # Seidel sums of the same population: permutation invariance, bit for bit
    sd = ctx.seidel_candidates(base, P["a"], P["h"])
    assert ctx.seidel_candidates(base[perm], P["a"], P["h"]).tobytes() == sd[perm].tobytes()
