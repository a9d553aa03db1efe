#!/bin/bash
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_resampler.py -m gpu -q -x > $O/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2u_pytest.log
tail -6 $O/r2u_pytest.log
timeout 200 python bench.py --workload resampler > $O/r2u_bench_resampler.json 2> $O/r2u_bench_resampler.err; echo "rc=$?"
tail -c 300 $O/r2u_bench_resampler.err
