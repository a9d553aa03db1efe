#!/bin/bash
# last GPU call of the round: the whole GPU tier at HEAD, then the north-star shape on one GPU (short)
set -x
O=gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > $O/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2x_pytest.log
tail -3 $O/r2x_pytest.log
timeout 120 python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline --no-profile > $O/r2x_c5_65536.json 2> $O/r2x_c5_65536.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*' $O/r2x_c5_65536.json
