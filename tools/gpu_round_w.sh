#!/bin/bash
# de-esser hand-offs gated by the rebuild mask (R_c1c gain stores, M_c2 gain loads, R_c3 coefficient prefetch)
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_full_length.py tests/test_gpu_workloads.py -m gpu -q -x > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2w_pytest.log
tail -4 $O/r2w_pytest.log
timeout 200 python bench.py --candidates 1024 --no-other-configs --no-cpu-baseline --no-profile > $O/r2w_c5_8192.json 2> $O/r2w_c5_8192.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*' $O/r2w_c5_8192.json
