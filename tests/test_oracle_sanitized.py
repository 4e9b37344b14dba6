"""SURVEY.md section 5 / VERDICT r1: the CPU oracle under -fsanitize=address,undefined.  compute-sanitizer is closed on the
GPU pool, so memory safety of the checker itself is shown here: every entry point the parity tests use, on ragged, empty,
one-element and hostile (NaN / Inf / miss / TIR) inputs, in a subprocess with libasan preloaded.  Any out-of-bounds access,
use of an uninitialised index or undefined arithmetic (signed overflow, bad shift, misaligned access) aborts the run."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import ort_b200 as ort
from oracle import oracle as orc, prelude as pre
P = ort.prescriptions
rng = np.random.default_rng(3)
for name in ("COOKE", "TESSAR", "SINGLET", "DOUBLE_GAUSS", "REFLECTIVE"):
    S = getattr(P, name)["surfaces"]
    a = getattr(P, name)["a"]
    so = pre.solve(S, a, getattr(P, name)["h"])
    for k_rays in (2, 3, 7, 64):
        for H in (0.0, 1.0):
            pre.full_trace(so, H, k_rays, threads=3)
    y0, x0 = rng.uniform(-30, 30, 257), rng.uniform(-30, 30, 257)          # many misses and TIR events
    u0, v0 = rng.uniform(-0.6, 0.6, 257), rng.uniform(-0.6, 0.6, 257)
    y0[:4] = [np.nan, np.inf, -np.inf, 0.0]
    orc.trace3d_batch(S, y0, x0, u0, v0, threads=2)
    orc.trace3d_ld_batch(S, y0, x0, u0, v0)
    orc.trace2d_batch(S, y0, u0)
    orc.trace3d_batch(S, y0[:0], x0[:0], u0[:0], v0[:0])                   # empty
    tau, phi, n = orc.lens(S)
    orc.paraxial_batch(tau, phi, y0, u0, a=a[:len(tau)] if len(a) >= len(tau) else None, clip=len(a) >= len(tau))
    orc.paraxial_trace(tau, phi, 1.0, 0.0, a=np.full(len(tau), 0.5), clip=True)
    orc.seidel(S, a, getattr(P, name)["h"])
    ext = np.vstack([S, [np.inf, 0.0, 1.0]])
    for ny, nx in ((0, 0), (1, 1), (5, 1), (1, 9), (13, 7)):
        orc.grid_trace(ext, rng.uniform(-25, 25, ny), rng.uniform(0, 25, nx), so.stop, 9.0, 1.0, u=0.05, threads=4)
        orc.grid_trace_ext(ext, rng.uniform(-25, 25, ny), rng.uniform(0, 25, nx), so.stop, 9.0, 1.0, u=0.05, a=np.append(a, np.inf),
                           rr=80.0, threads=4)
Sp = P.PARABOLA["surfaces"]
orc.trace2d_batch(Sp[:, :3], np.array([30.0, 1.0]), np.zeros(2), K=Sp[:, 3], aspheric=True)
R = P.perturbed_triplets(9)
ext = np.concatenate([R, np.tile(np.array([np.inf, 0.0, 1.0, 0.0])[None, :, None], (9, 1, 1))], axis=2)
ext[:, 1, -2] = 77.0
orc.candidates(ext, np.linspace(-12, 12, 9), np.linspace(0, 12, 5), 5, 10.3, 21.0, 0.1)
print("SANITIZED-OK")
'''


def test_oracle_runs_clean_under_asan_ubsan():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "san"])
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1",
               ORT_ORACLE_LIB=os.path.join(ROOT, "oracle", "libort_oracle_san.so"), OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SANITIZED-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
