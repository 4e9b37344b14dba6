"""The communicator inside libort_b200.so (include/ort_b200.h, "communicator inside the library"): NCCL resolved with dlopen,
ort_opts.gather_stats (ncclAllGather + rank-order merge kernel behind the sweep), ort_trace3d_grid_multi (one process, n
GPUs), ort_candidates_sharded (BASELINE config 5 across ranks).  Tests that need two GPUs skip on a one-GPU box; the
one-GPU tests still run the NCCL path (world 1) and the merge kernel (records staged by hand)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _cooke_inputs(ort, ctx, Hs=(0.0, 0.7, 1.0), ny=64, nx=32):
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=ctx)
    p = ort.host._full_trace_setup(s.layout, s, list(Hs), 64, None, ctx)
    ys = np.stack([ort.host.jl_range(p["y1"][j], p["y2"][j], ny) for j in range(len(Hs))])
    xs = ort.host.jl_range(0.0, p["y_EP"], nx)
    flds = [dict(u=float(p["u"][j]), h_prime=float(p["h_prime"][j])) for j in range(len(Hs))]
    return s, p, ys, xs, flds


def test_merge_kernel_equals_host_merge(ctx, ort):
    """k_merge_stats (the kernel behind opts.gather_stats) folds records in rank order with the operations of the host
    function ort_merge_stats: bit-identical, on 4 hand-made shards of a real sweep (one of them empty)"""
    import torch
    s, p, ys, xs, flds = _cooke_inputs(ort, ctx)
    ctx.set_layout(p["ext"], p["K"])
    shards = [(0, 20), (20, 20), (20, 47), (47, 64)]
    recs = np.stack([ctx.trace3d_grid(flds, ys[:, lo:hi], xs, p["stop"], p["a_stop"], want=("stats",))["stats"] for lo, hi in shards])
    host = ort._lib.merge_stats_c(recs)
    d_in = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint8).reshape(-1)).cuda()
    d_out = torch.zeros(len(flds) * ort.STATS_BYTES, dtype=torch.uint8, device="cuda")
    ctx.merge_stats_dev(d_in.data_ptr(), len(shards), len(flds), d_out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dev = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=ort.STATS_DTYPE)
    assert dev.tobytes() == host.tobytes()
    whole = ctx.trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"], want=("stats",))["stats"]
    for f in range(len(flds)):
        assert dev[f]["n_kept"] == whole[f]["n_kept"] and dev[f]["r_max"] == whole[f]["r_max"]
        assert abs(ort.rms_from_stats(dev[f]) - ort.rms_from_stats(whole[f])) < 1e-13


def test_world1_communicator_gather_is_identity(ort):
    """NCCL loads (dlopen), a 1-rank communicator initialises, and a sweep with opts.gather_stats returns merged == local
    == the plain sweep, host- and device-pointer paths; without a communicator the option is refused with ORT_ENCCL"""
    import torch
    c = ort.Context(0)
    try:
        s, p, ys, xs, flds = _cooke_inputs(ort, c)
        c.set_layout(p["ext"], p["K"])
        with pytest.raises(ort.OrtError) as ei:
            c.trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"], gather=True)
        assert ei.value.code == ort._lib.ORT_ENCCL
        assert c.comm_info()["world"] == 0
        c.comm_init_rank(ort._lib.comm_unique_id(), 0, 1)
        info = c.comm_info()
        assert info["world"] == 1 and info["rank"] == 0 and info["nccl_version"] >= 21800
        plain = c.trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"], compact=True)
        g = c.trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"], compact=True, gather=True)
        assert g["stats"].tobytes() == g["stats_local"].tobytes() == plain["stats"].tobytes()
        assert np.array_equal(g["mask"], plain["mask"])
        # device-pointer form
        nf, NN = len(flds), ys.shape[1] * len(xs)
        d_ys, d_xs = torch.from_numpy(ys).cuda(), torch.from_numpy(xs).cuda()
        d_m = torch.zeros(nf * ort.STATS_BYTES, dtype=torch.uint8, device="cuda")
        d_l = torch.zeros_like(d_m)
        d_mask = torch.zeros(nf * NN, dtype=torch.uint8, device="cuda")
        c.trace3d_grid_dev(flds, d_ys.data_ptr(), ys.shape[1], d_xs.data_ptr(), len(xs), p["stop"], p["a_stop"],
                           dict(mask=d_mask.data_ptr(), stats=d_m.data_ptr(), stats_local=d_l.data_ptr()),
                           stream=torch.cuda.current_stream().cuda_stream, ys_per_field=True, gather=True)
        torch.cuda.synchronize()
        assert d_m.cpu().numpy().tobytes() == d_l.cpu().numpy().tobytes() == plain["stats"].tobytes()
        # config 5 over a 1-rank communicator == the plain prelude + aimed sweep
        Pq = ort.prescriptions.COOKE
        R = ort.prescriptions.perturbed_triplets(257)
        tab, aim = c.candidates_sharded(R, Pq["a"], Pq["h"], 0.7, 32, want_aim=True)
        aim0 = c.aim_candidates(R, Pq["a"], Pq["h"], 0.7)
        tab0 = c.trace3d_candidates_aimed(R, aim0, 32, 16)
        assert np.array_equal(aim, aim0, equal_nan=True) and np.array_equal(tab, tab0, equal_nan=True)
        c.comm_free()
        assert c.comm_info()["world"] == 0
    finally:
        c.close()


def test_grid_multi_single_context_equals_plain(ctx, ort):
    """ort_trace3d_grid_multi over one context is ort_trace3d_grid (no communicator needed), full and compacted"""
    s, p, ys, xs, flds = _cooke_inputs(ort, ctx, ny=53, nx=29)
    ctx.set_layout(p["ext"], p["K"])
    want = ("ex", "ey", "r", "theta", "mask", "flags", "stats")
    for compact in (False, True):
        a = ctx.trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"], compact=compact, want=want)
        b = ort._lib.trace3d_grid_multi([ctx], flds, ys, xs, p["stop"], p["a_stop"], compact=compact, want=want)
        assert a["stats"].tobytes() == b["stats"].tobytes() == b["stats_local"][0].tobytes()
        for f in range(len(flds)):
            n = int(a["stats"][f]["n_kept"]) if compact else ys.shape[1] * len(xs)
            for k in ("ex", "ey", "r", "theta"):
                assert np.array_equal(a[k][f][:n], b[k][f][:n], equal_nan=True), (compact, f, k)
        assert np.array_equal(a["mask"], b["mask"]) and np.array_equal(a["flags"], b["flags"])


def test_streams_share_scratch_safely(ctx, ort):
    """two *_dev sweeps with compaction of one context on two different streams share the context's scratch
    slots; the library orders them itself, so both results equal the serial ones"""
    import torch
    s, p, ys, xs, flds = _cooke_inputs(ort, ctx, ny=640, nx=320)
    ctx.set_layout(p["ext"], p["K"])
    nf, ny, nx = len(flds), ys.shape[1], len(xs)
    NN = ny * nx
    d_ys, d_xs = torch.from_numpy(ys).cuda(), torch.from_numpy(xs).cuda()
    d_ys2 = torch.from_numpy(np.ascontiguousarray(ys[:, ::-1])).cuda()          # a different sweep: rows reversed
    outs = []
    for _ in range(2):
        outs.append({k: torch.zeros(nf * NN, dtype=torch.float64, device="cuda") for k in ("ex", "ey")})
        outs[-1]["stats"] = torch.zeros(nf * ort.STATS_BYTES, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(o, ysd, stream):
        ctx.trace3d_grid_dev(flds, ysd.data_ptr(), ny, d_xs.data_ptr(), nx, p["stop"], p["a_stop"],
                             {k: t.data_ptr() for k, t in o.items()}, stream=stream.cuda_stream, compact=True, ys_per_field=True)
    torch.cuda.synchronize()
    run(outs[0], d_ys, s1)
    torch.cuda.synchronize()
    ref0 = {k: t.clone() for k, t in outs[0].items()}
    run(outs[1], d_ys2, s1)
    torch.cuda.synchronize()
    ref1 = {k: t.clone() for k, t in outs[1].items()}
    for t in list(outs[0].values()) + list(outs[1].values()):
        t.zero_()
    torch.cuda.synchronize()
    for _ in range(3):                       # back to back on two streams, no host synchronisation in between
        run(outs[0], d_ys, s1)
        run(outs[1], d_ys2, s2)
    torch.cuda.synchronize()
    for k in ("ex", "ey", "stats"):
        assert torch.equal(outs[0][k], ref0[k]) and torch.equal(outs[1][k], ref1[k]), k


@pytest.mark.skipif("_ngpu() < 2")
def test_grid_multi_two_gpus_equals_one(ort):
    """one process, two contexts: ort_comm_init_all + ort_trace3d_grid_multi.  Rows are block-sharded, so every ray-level
    output is bit-identical to the one-GPU sweep in the reference's order (full and compacted); the merged statistics
    equal the host merge of the two shard records bit for bit"""
    cs = [ort.Context(0), ort.Context(1)]
    try:
        s, p, ys, xs, flds = _cooke_inputs(ort, cs[0], ny=129, nx=67)
        ort._lib.comm_init_all(cs)
        cs[0].set_layout(p["ext"], p["K"])
        want = ("ex", "ey", "r", "theta", "mask", "flags", "stats")
        for arith in (ort.FAST, ort.STRICT):
            for compact in (False, True):
                one = cs[0].trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"], compact=compact, want=want, arith=arith)
                two = ort._lib.trace3d_grid_multi(cs, flds, ys, xs, p["stop"], p["a_stop"], compact=compact, want=want, arith=arith)
                assert np.array_equal(one["mask"], two["mask"]) and np.array_equal(one["flags"], two["flags"])
                assert two["stats"].tobytes() == ort._lib.merge_stats_c(two["stats_local"]).tobytes()
                for f in range(len(flds)):
                    assert int(two["stats"][f]["n_kept"]) == int(one["stats"][f]["n_kept"])
                    assert two["stats"][f]["r_max"] == one["stats"][f]["r_max"]
                    assert abs(ort.rms_from_stats(two["stats"][f]) - ort.rms_from_stats(one["stats"][f])) < 1e-13
                    n = int(one["stats"][f]["n_kept"]) if compact else ys.shape[1] * len(xs)
                    for k in ("ex", "ey", "r", "theta"):
                        assert np.array_equal(one[k][f][:n], two[k][f][:n], equal_nan=True), (arith, compact, f, k)
        # polynomial terms travel with the layout of ctxs[0]: the device table (coefficients and k c_k) is copied to the
        # other device, short rows go as a kernel parameter; 9 and 13 coefficient columns
        rows = p["ext"].shape[0]
        for ncol in (9, 13):
            P = np.zeros((rows, ncol)); P[1, 4] = 2e-7; P[3, 6] = -1e-10; P[2, ncol - 1] = 1e-13 if ncol == 9 else 3e-19
            cs[0].set_layout(p["ext"], p["K"]); cs[0].set_polynomials(P)
            for arith in (ort.FAST, ort.STRICT):
                one = cs[0].trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"], want=want, arith=arith)
                two = ort._lib.trace3d_grid_multi(cs, flds, ys, xs, p["stop"], p["a_stop"], want=want, arith=arith)
                assert np.array_equal(one["mask"], two["mask"]) and np.array_equal(one["flags"], two["flags"])
                for k in ("ex", "ey", "r", "theta"):
                    assert np.array_equal(one[k], two[k], equal_nan=True), (ncol, arith, k)
        cs[0].set_polynomials(None)
    finally:
        for c in cs:
            c.close()


def _build_example(ort, tmp_path):
    exe = str(tmp_path / "sharded_sweep")
    libdir = os.path.dirname(ort._lib.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "sharded_sweep.c"), "-o", exe, "-L", libdir, "-lort_b200", "-lm",
                           "-Wl,-rpath," + libdir])
    return exe


def _kv(line):
    return dict(item.split("=") for item in line.split() if "=" in item)


def test_plain_c_sharded_example_one_gpu(ctx, ort, tmp_path):
    """examples/sharded_sweep.c from plain C on one GPU: rank form with world 1 (NCCL inside the library) and multi form"""
    exe = _build_example(ort, tmp_path)
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=ctx)
    p = ort.host._full_trace_setup(s.layout, s, [0.7], 64, None, ctx)
    args = [repr(float(v)) for v in (p["y1"][0], p["y2"][0], p["y_EP"], p["u"][0], p["h_prime"][0], p["focus"])]
    e = ort.full_trace(s, 0.7)
    o1 = _kv(subprocess.run([exe, "rank", "0", "1", str(tmp_path / "id1")] + args, capture_output=True, text=True, check=True).stdout)
    o2 = _kv(subprocess.run([exe, "multi", "1"] + args, capture_output=True, text=True, check=True).stdout)
    for o in (o1, o2):
        assert int(o["n_kept"]) == len(e.x) // 2 and abs(float(o["rms"]) - e.RMS) < 1e-12 * 25
    assert int(o1["nccl"]) >= 21800 and o1["rank"] == "0/1"


@pytest.mark.skipif("_ngpu() < 2")
def test_plain_c_two_ranks(ctx, ort, tmp_path):
    """a plain-C program, one process per GPU, two ranks: every rank prints the same merged statistics,
    equal to the one-GPU sweep; the single-process multi form agrees too"""
    exe = _build_example(ort, tmp_path)
    P = ort.prescriptions.COOKE
    s = ort.solve(P["surfaces"], P["a"], P["h"], backend=ctx)
    p = ort.host._full_trace_setup(s.layout, s, [0.7], 64, None, ctx)
    args = [repr(float(v)) for v in (p["y1"][0], p["y2"][0], p["y_EP"], p["u"][0], p["h_prime"][0], p["focus"])]
    e = ort.full_trace(s, 0.7)
    idf = str(tmp_path / "id2")
    procs = [subprocess.Popen([exe, "rank", str(r), "2", idf] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = []
    for pr in procs:
        so, se = pr.communicate(timeout=240)
        assert pr.returncode == 0, se
        outs.append(_kv(so))
    assert outs[0]["rms"] == outs[1]["rms"] and outs[0]["mean_y"] == outs[1]["mean_y"]          # identical text = identical bits
    assert int(outs[0]["n_kept"]) == int(outs[1]["n_kept"]) == len(e.x) // 2
    assert int(outs[0]["n_local"]) + int(outs[1]["n_local"]) == int(outs[0]["n_kept"])
    assert abs(float(outs[0]["rms"]) - e.RMS) < 1e-12 * 25
    m = _kv(subprocess.run([exe, "multi", "2"] + args, capture_output=True, text=True, check=True).stdout)
    assert m["rms"] == outs[0]["rms"] and int(m["n_kept"]) == int(m["sum_local"]) == len(e.x) // 2


def _rank_worker(rank, world, idbytes, q):
    sys.path.insert(0, ROOT)
    import ort_b200 as ort
    c = ort.Context(rank)
    c.comm_init_rank(idbytes, rank, world)
    s, p, ys, xs, flds = _cooke_inputs(ort, c, ny=129, nx=67)
    c.set_layout(p["ext"], p["K"])
    lo, hi = ort._lib.comm_range(ys.shape[1], rank, world)
    g = c.trace3d_grid(flds, ys[:, lo:hi], xs, p["stop"], p["a_stop"], compact=True, gather=True)
    Pq = ort.prescriptions.COOKE
    R = ort.prescriptions.perturbed_triplets(1001)                     # uneven ranges: the grouped-broadcast path
    tab = c.candidates_sharded(R, Pq["a"], Pq["h"], 0.7, 32)
    tab_even = c.candidates_sharded(R[:1000], Pq["a"], Pq["h"], 0.7, 32)   # even ranges: the in-place all-gather
    q.put((rank, g["stats"].tobytes(), g["stats_local"].tobytes(), tab.tobytes(), tab_even.tobytes()))
    c.close()


@pytest.mark.skipif("_ngpu() < 2")
def test_two_rank_gather_and_sharded_candidates(ctx, ort):
    """one process per GPU through the Python binding: merged statistics identical on both ranks and equal to the host
    merge of the two shard records; the sharded population sweep returns the complete merit table on every rank, equal
    to the one-GPU table bit for bit (candidates are independent)"""
    import torch.multiprocessing as mp
    idb = ort._lib.comm_unique_id()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_rank_worker, args=(r, 2, idb, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0][1] == res[1][1]                                       # merged records: same bytes on both ranks
    loc = np.stack([np.frombuffer(r[2], dtype=ort.STATS_DTYPE) for r in res])
    assert ort._lib.merge_stats_c(loc).tobytes() == res[0][1]
    s, p, ys, xs, flds = _cooke_inputs(ort, ctx, ny=129, nx=67)
    ctx.set_layout(p["ext"], p["K"])
    whole = ctx.trace3d_grid(flds, ys, xs, p["stop"], p["a_stop"])["stats"]
    merged = np.frombuffer(res[0][1], dtype=ort.STATS_DTYPE)
    assert [int(x) for x in merged["n_kept"]] == [int(x) for x in whole["n_kept"]]
    Pq = ort.prescriptions.COOKE
    R = ort.prescriptions.perturbed_triplets(1001)
    one = ctx.trace3d_candidates_aimed(R, ctx.aim_candidates(R, Pq["a"], Pq["h"], 0.7), 32, 16)
    assert res[0][3] == res[1][3] == one.tobytes()
    assert res[0][4] == res[1][4] == one[:1000].tobytes()
