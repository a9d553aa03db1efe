#!/bin/bash
# product resampler simulator on one B200: GPU parity tier, the bench sub-record, launch list + ncu --set full of k_resample
set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_resampler.py -m gpu -q -x > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2m_pytest.log
tail -15 $O/r2m_pytest.log
timeout 300 python bench.py --workload resampler > $O/r2m_bench_resampler.json 2> $O/r2m_bench_resampler.err; echo "bench rc=$?"
tail -c 300 $O/r2m_bench_resampler.err
cat $O/r2m_bench_resampler.json | head -c 3000
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name k_resample --launch-skip 3 -c 2 -o $O/r2m_resample \
  python bench.py --workload resampler --steps 1 --warmup 3 --no-cpu-baseline > $O/r2m_ncu.log 2>&1
du -sh $O
