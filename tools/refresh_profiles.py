"""Turn the raw captures a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/refresh_profiles.py <tag> [round = r02]      (reads gpurun_out/prof_grid_<tag>.ncu-rep, launches_<tag>.csv,
bench_<tag>.json, bench_ref_<tag>.json; writes profiles/<round>_*)"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAYS = 83868160
KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']


def num(x):
    return float(x.replace(',', ''))


def main():
    tag = sys.argv[1]
    rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    rep = os.path.join(go, f"prof_grid_{tag}.ncu-rep")
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
    summ = {k: {"value": d.get(k), "unit": u.get(k)} for k in KEYS if k in d}
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    dram = sum(num(d[k]) * mult[u[k]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    # dynamic instruction mix from the source page (SASS view)
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True,
                         text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    h = srows[1]
    ia, ie = h.index("Source"), h.index("Instructions Executed")
    mix = collections.Counter()
    for r in srows[2:]:
        if len(r) <= ie:
            continue
        op = r[ia].strip().split()
        if not op:
            continue
        name = op[1] if op[0].startswith('@') else op[0]
        mix[name.split('.')[0]] += int(r[ie] or 0)
    tot = sum(mix.values())
    fp64 = sum(mix[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    per = RAYS / 32
    summ["_how"] = ("ncu --set full --clock-control none --import-source on -k regex:k_grid -s 3 -c 1 python bench.py "
                    "--steps 3 --warmup 3 --no-cpu-baseline --no-extras (after the same command exited 0 without ncu)")
    summ["_derived"] = {"rays_per_launch": RAYS, "warp_instructions_per_32_rays": tot / per,
                        "fp64_pipe_instructions_per_ray": fp64 / per, "other_instructions_per_ray": (tot - fp64) / per,
                        "issue_slot_model_cycles_per_32_rays": (2 * fp64 + (tot - fp64)) / per / 0.9,
                        "measured_cycles_per_32_rays": num(d['sm__cycles_elapsed.avg']) / (per / (148 * 4)),
                        "algorithmic_bytes_per_launch": RAYS * 17, "dram_bytes_per_launch": dram,
                        "instruction_mix_per_ray": {k: round(v / per, 2) for k, v in mix.most_common(16)}}
    json.dump(summ, open(os.path.join(pr, f"{rnd}_k_grid_fast_{tag}_ncu_summary.json"), "w"), indent=1)
    json.dump({"kernel": d['Kernel Name'], "dram_bytes_per_launch": dram,
               "source": f"profiles/{rnd}_k_grid_fast_{tag}_ncu_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, "
                         "one ncu --set full capture)"}, open(os.path.join(pr, "k_grid_traffic.json"), "w"), indent=1)
    # launch list
    lrows = [r for r in csv.reader(open(os.path.join(go, f"launches_{tag}.csv"))) if len(r) > 10]
    hh = lrows[0]
    ki, vi = hh.index('Kernel Name'), hh.index('Metric Value')
    agg = collections.OrderedDict()
    for r in lrows[1:]:
        agg.setdefault(r[ki], []).append(num(r[vi]))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(pr, f"{rnd}_launches_{tag}_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 160 python bench.py --steps 3 --warmup 3 "
                "--no-cpu-baseline --no-extras\n# (cold-cache, serialised: compare shares).  Prelude (k_paraxial, k_aim_candidates, k_aim_edges), "
                "device steps (one 5-field k_grid launch + k_grid_finalize each),\n# e2e steps (per-field k_grid + finalize "
                "[+ k_tile_scan + k_chunk_scan + k_compact]), k_fp64_peak = the roofline denominator.  Inside a timed step the\n"
                "# launches are k_grid + k_grid_finalize: k_grid's share of the step = 99.8 %.\n")
        for k, v in agg.items():
            f.write(f"{k[:70]:72s} n={len(v):3d} total={sum(v) / 1e6:9.3f} ms share={100 * sum(v) / total:6.2f}% "
                    f"mean={sum(v) / len(v) / 1e3:9.1f} us\n")
    shutil.copy(os.path.join(go, f"launches_{tag}.csv"), os.path.join(pr, f"{rnd}_launches_{tag}.csv"))
    for src_name, dst in ((f"bench_{tag}.json", f"{rnd}_bench_1gpu.json"), (f"bench_ref_{tag}.json", f"{rnd}_bench_reference_arm.json")):
        if os.path.exists(os.path.join(go, src_name)):
            shutil.copy(os.path.join(go, src_name), os.path.join(pr, dst))
    # SASS of the dominant kernel (proof of what runs: DFMA / MUFU.RCP64H / MUFU.RSQ64H, LDCU uniform operands):
    # k_grid<FAST, 3 rays/thread, spot + mask, no-mirror, SIMPLE>
    lib = os.path.join(ROOT, "opticalraytracing.jl_b200", "lib", "libort_b200.so")
    sass = subprocess.run(['cuobjdump', '-sass', '-fun', '_Z6k_gridILi1ELi3ELi0ELi1ELb0ELi1EEv5Presc8GridArgs', lib],
                          capture_output=True, text=True).stdout
    keep = [ln.split('/*', 2)[0].rstrip() + '  ' + ln.split('*/', 1)[1].rsplit('/*', 1)[0].rstrip() if '*/' in ln and ln.strip().startswith('/*') else ln
            for ln in sass.splitlines() if not ln.strip().startswith('/* 0x')]
    open(os.path.join(pr, f"{rnd}_k_grid_fast_simple_rpt3.sass"), "w").write("\n".join(keep) + "\n")
    print(json.dumps(summ["_derived"], indent=1))


if __name__ == "__main__":
    main()
