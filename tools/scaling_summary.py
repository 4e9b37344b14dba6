"""profiles/<round>_bench_{1,2,4,8}gpu.json -> profiles/<round>_scaling_summary.json (weak scaling of the headline, host-pointer
e2e, the strong-scaling records).  usage: python tools/scaling_summary.py [round = r02]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = []
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "profiles", f"{rnd}_bench_{n}gpu.json")
    if not os.path.exists(p):
        continue
    b = json.load(open(p))
    st = b["config"]["strong"]
    c3, c5 = st["config3_1e9_ray_sweep"], st["config5_65536_candidates"]
    rows.append({"n": n, "value": b["value"], "ms_per_step": b["ms_per_step"], "e2e": b["e2e"]["value"], "e2e_ms": b["e2e"]["ms_per_step"],
                 "config3_ms": c3["ms"], "config5_ms": c5["ms"], "config3_rms": c3["rms_mm"],
                 "config5_checksum": c5.get("table_checksum")})
r1 = rows[0]
for r in rows:
    r["weak_eff"] = r["value"] / (r["n"] * r1["value"])
    r["e2e_eff"] = r["e2e"] / (r["n"] * r1["e2e"])
    r["config3_strong_eff"] = r1["config3_ms"] / (r["n"] * r["config3_ms"])
    r["config5_strong_eff"] = r1["config5_ms"] / (r["n"] * r["config5_ms"])
out = {"_what": f"bench.py at N = 1, 2, 4, 8 (profiles/{rnd}_bench_{{N}}gpu.json): weak scaling of the headline, host-pointer e2e, "
                "and the strong-scaling records (fixed total work)", "rows": rows}
json.dump(out, open(os.path.join(ROOT, "profiles", f"{rnd}_scaling_summary.json"), "w"), indent=1)
for r in rows:
    print(r["n"], f'{r["value"]:.4g}', f'{r["ms_per_step"]:.4f} ms', f'weak {r["weak_eff"]:.4f}', f'e2e {r["e2e"]:.4g} ({r["e2e_eff"]:.3f})',
          f'c3 {r["config3_ms"]:.3f} ms ({r["config3_strong_eff"]:.3f})', f'c5 {r["config5_ms"]:.4f} ms ({r["config5_strong_eff"]:.3f})')
