#!/bin/bash
# resampler kernel with the mirrored phase table: GPU parity tier, groups of 8 against groups of 4
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_resampler.py -m gpu -q -x > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s_pytest.log
tail -6 $O/r2s_pytest.log
timeout 200 python bench.py --workload resampler --no-cpu-baseline > $O/r2s_bench_g8.json 2> $O/r2s_bench_g8.err; echo "rc=$?"
AFSIM_RESAMPLE_GROUP=4 timeout 200 python bench.py --workload resampler --no-cpu-baseline > $O/r2s_bench_g4.json 2> $O/r2s_bench_g4.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*' $O/r2s_bench_g*.json
