"""BASELINE config 5 across GPUs: 65 536 perturbed triplet prescriptions x 4096 rays, candidates sharded by contiguous
ranges over the ranks (one process per GPU, torchrun), per-candidate prelude + aimed sweep on device-resident
prescriptions, ONE NCCL all-gather of the 32 B-per-candidate merit table.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config5_multi.py [C] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import ort_b200 as ort


def main():
    C = int(float(sys.argv[1])) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    real_stdout = os.dup(1); os.dup2(2, 1)                  # NCCL banners must not reach stdout
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = ort.Context(local)
    P = ort.prescriptions.COOKE
    base = ort.prescriptions.perturbed_triplets(C)
    lo, hi = ort.distributed.shard_rows(C, rank, world)
    n_loc = hi - lo
    per = -(-C // world)
    d_R = torch.from_numpy(base[lo:hi].copy()).to(dev)
    d_aim = torch.empty((n_loc, ort._lib.AIM_NOUT), dtype=torch.float64, device=dev)
    d_out = torch.full((per, 4), float("nan"), dtype=torch.float64, device=dev)
    d_all = torch.empty((world, per, 4), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.aim_candidates_dev(8, n_loc, d_R.data_ptr(), P["a"], P["h"], 0.7, d_aim.data_ptr(), stream=st)
        ctx.trace3d_candidates_aimed_dev(8, n_loc, d_R.data_ptr(), d_aim.data_ptr(), 64, 64, d_out.data_ptr(), stream=st)
        if world > 1:
            dist.all_gather_into_tensor(d_all.view(-1), d_out.view(-1))

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    table = (d_all if world > 1 else d_out[None]).cpu().numpy()
    rows = np.concatenate([table[r, :ort.distributed.shard_rows(C, r, world)[1] - ort.distributed.shard_rows(C, r, world)[0]]
                           for r in range(world)])
    if rank == 0:
        line = {"config": "BASELINE config 5", "n_gpus": world, "candidates": C, "rays_per_candidate": 4096, "ms_per_step": ms,
                "candidates_per_s": C / ms * 1e3, "rays_per_s": C * 4096 / ms * 1e3,
                "intersections_per_s_x7": C * 4096 * 7 / ms * 1e3, "steps": steps,
                "rms_range_mm": [float(np.nanmin(rows[:, 3])), float(np.nanmax(rows[:, 3]))],
                "best_candidate": int(np.nanargmin(rows[:, 3])), "failed": int(np.isnan(rows[:, 3]).sum()),
                "exchange": f"all_gather_into_tensor of {per * 32} B per rank" if world > 1 else "none"}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
