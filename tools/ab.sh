#!/bin/bash
# A/B: run bench.py against several builds of the library (ORT_B200_LIB)
P='import json,sys; d=json.load(sys.stdin); print(sys.argv[1], "ms_step=%.4f kern_ms=%.4f frac=%.4f value=%.4g" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]))'
for v in "$@"; do
  ORT_B200_LIB=$PWD/opticalraytracing.jl_b200/lib/libort_b200_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline $ABFLAGS 2>>gpurun_out/ab.err | python -c "$P" $v
done
