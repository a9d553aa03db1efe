#!/bin/bash
set -x
O=gpurun_out
timeout 200 python bench.py --workload resampler > $O/r2t_bench_resampler.json 2> $O/r2t_bench_resampler.err && \
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_resample --launch-skip 3 -c 1 -o $O/r2t_resample \
  python bench.py --workload resampler --steps 1 --warmup 3 --no-cpu-baseline > $O/r2t_ncu.log 2>&1
echo rc=$?
