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
