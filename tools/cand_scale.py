import sys, json, numpy as np, torch
sys.path.insert(0, '.')
import ort_b200 as ort
C = 32768
dev = torch.device("cuda", 0); ctx = ort.Context(0); st = torch.cuda.current_stream().cuda_stream
Pq = ort.prescriptions.COOKE
base = ort.prescriptions.perturbed_triplets(C)
d_R = torch.from_numpy(base).to(dev); d_aim = torch.empty((C, 24), dtype=torch.float64, device=dev); d_out = torch.empty((C, 4), dtype=torch.float64, device=dev)
ctx.aim_candidates_dev(8, C, d_R.data_ptr(), Pq["a"], Pq["h"], 0.7, d_aim.data_ptr(), stream=st)
res = {}
for ny, nx in ((16, 16), (32, 32), (64, 32), (64, 64), (128, 64), (128, 128)):
    f = lambda: ctx.trace3d_candidates_aimed_dev(8, C, d_R.data_ptr(), d_aim.data_ptr(), ny, nx, d_out.data_ptr(), arith=ort.FAST, stream=st)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[f"{ny}x{nx}"] = {"ms": ms, "ns_per_candidate": ms * 1e6 / C, "ps_per_ray": ms * 1e9 / C / (ny * nx)}
print(json.dumps(res, indent=0))
