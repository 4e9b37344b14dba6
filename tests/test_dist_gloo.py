"""The N > 1 path on CPU: world_size-2 gloo, y-rows sharded across ranks, one all-gather of the
per-field statistics records, ranks merged in rank order.  Compute is the oracle-backed test
backend; the sharding / collective / merge code is the product's (ort_b200.distributed)."""
import math
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, HERE]
    import torch.distributed as dist
    import ort_b200 as ort
    from oracle import prelude as pre
    from oracle_backend import OracleBackend
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    P = ort.prescriptions.COOKE
    so = pre.solve(P["surfaces"], P["a"], P["h"])
    ps = [pre.full_trace_inputs(so, H, 64) for H in (0.0, 0.7, 1.0)]
    be = OracleBackend()
    be.set_layout(ps[0].ext, ps[0].K)
    ys = np.stack([p.ys for p in ps])
    fields = [dict(u=p.u, v=p.v, h_prime=p.h_prime) for p in ps]
    res, merged, rms = ort.distributed.sharded_sweep(be, fields, ys, ps[0].xs, ps[0].stop, ps[0].a_stop, rank, world,
                                                     want=("ex", "ey", "r", "theta", "mask", "stats"), compact=True)
    lo, hi = ort.distributed.shard_rows(64, rank, world)
    q.put((rank, rms, [int(m["n_kept"]) for m in merged], [float(m["r_max"]) for m in merged],
           [res["ex"][f][:int(res["stats"][f]["n_kept"])].copy() for f in range(3)], (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_sweep_matches_single():
    sys.path[:0] = [ROOT, HERE]
    import ort_b200 as ort
    from oracle import oracle as orc, prelude as pre
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = ort.prescriptions.COOKE
    so = pre.solve(P["surfaces"], P["a"], P["h"])
    assert out[0][5] == (0, 32) and out[1][5] == (32, 64)
    for f, H in enumerate((0.0, 0.7, 1.0)):
        ref = pre.full_trace(so, H, 64)
        n = len(ref.x) // 2
        # identical on every rank (merged in rank order), and equal to the single-process result
        assert out[0][1][f] == out[1][1][f]
        assert math.isclose(out[0][1][f], ref.RMS, rel_tol=1e-12)
        assert out[0][2][f] == n == out[1][2][f]
        # concatenating the ranks' compacted outputs in rank order reproduces the reference's push! order
        cat = np.concatenate([out[0][4][f], out[1][4][f]])
        assert np.array_equal(cat, ref.x[:n])


def _cand_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, HERE]
    import torch.distributed as dist
    import ort_b200 as ort
    from oracle_backend import OracleBackend
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    P = ort.prescriptions.COOKE
    RtnK = ort.prescriptions.perturbed_triplets(5)            # 5 candidates over 2 ranks: ragged ranges 3 + 2
    table, rng = ort.distributed.sharded_candidates(OracleBackend(), RtnK, P["a"], P["h"], 0.7, 32, rank, world)
    q.put((rank, table, rng))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_candidates():
    """config 5 over 2 ranks: contiguous candidate ranges, no data-path collective, one all-gather of the merit table;
    every rank ends with the same table as a single process."""
    sys.path[:0] = [ROOT, HERE]
    import ort_b200 as ort
    from oracle_backend import OracleBackend
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cand_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = ort.prescriptions.COOKE
    single, _ = ort.distributed.sharded_candidates(OracleBackend(), ort.prescriptions.perturbed_triplets(5), P["a"], P["h"],
                                                   0.7, 32)
    assert out[0][2] == (0, 3) and out[1][2] == (3, 5)
    assert out[0][1].tobytes() == out[1][1].tobytes() == single.tobytes()
