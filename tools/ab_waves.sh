#!/bin/bash
# kernel time of the bench sweep against the number of CTA waves per launch (ORT_GRID_WAVES)
P='import json,sys; d=json.load(sys.stdin); print("waves", sys.argv[1], "ms_step=%.4f kern_ms=%.4f frac=%.4f rms=%s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["config"]["spot_rms_mm"][:2]))'
for w in "$@"; do
  ORT_GRID_WAVES=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/ab.err | python -c "$P" $w
done
