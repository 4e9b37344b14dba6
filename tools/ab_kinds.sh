#!/bin/bash
# A/B of library builds (ORT_B200_LIB): bench.py headline + tools/bench_kinds.py (every surface body)
P='import json,sys; d=json.load(sys.stdin); print(sys.argv[1], "ms_step=%.4f kern_ms=%.4f frac=%.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))'
for v in "$@"; do
  export ORT_B200_LIB=$PWD/opticalraytracing.jl_b200/lib/libort_b200_$v.so
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/ab.err | python -c "$P" $v
  echo "$v kinds $(python tools/bench_kinds.py 2>>gpurun_out/ab.err)"
done
