#!/usr/bin/env python
"""K3 (paraxial y-nu trace, 40-row Lens) on N device-resident rays: plain FAST, FAST + clip (a = 25: ~85 % of the rays clip),
FAST + clip with an aperture nothing reaches, STRICT.  CUDA-event timing via ort_profile_*.  ORT_B200_LIB selects the build.
usage: bench_paraxial.py [N]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ort_b200 as ort  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6544.0
dev = torch.device("cuda", 0)
ctx = ort.Context(0)
st = torch.cuda.current_stream().cuda_stream
lens = ort.make_lens(ort.prescriptions.zoom20())
g = torch.Generator(device=dev); g.manual_seed(42)
y0 = torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 20 - 10
w0 = torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 0.4 - 0.2
y, w = torch.empty_like(y0), torch.empty_like(w0)
ci = torch.empty(N, dtype=torch.int32, device=dev)
out = {"N": N}
for name, arith, clip, aval in (("fast", ort.FAST, False, None), ("fast_clip_a25", ort.FAST, True, 25.0), ("fast_clip_never", ort.FAST, True, 1e6),
                                ("strict", ort.STRICT, False, None), ("strict_clip_a25", ort.STRICT, True, 25.0)):
    a = None if aval is None else np.full(len(lens.tau), aval)

    def run():
        ctx.paraxial_batch_dev(lens.tau, lens.phi, N, y0.data_ptr(), w0.data_ptr(), y.data_ptr(), w.data_ptr(),
                               ci.data_ptr() if clip else None, a=a, clip=clip, arith=arith, stream=st)
    run(); run()
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    ms = float(np.mean(ctx.profile_read()))
    ctx.profile_enable(False)
    byts = N * (36 if clip else 32)
    out[name] = {"ms": round(ms, 4), "GBps": round(byts / ms / 1e6, 1), "hbm_frac": round(byts / ms / 1e6 / HBM, 4)}
    if clip:
        out[name]["clipped_frac"] = round(float((ci[:1 << 24] != 0).double().mean()), 4)
print(json.dumps(out))
