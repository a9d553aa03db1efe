#!/bin/bash
# round-2 GPU pass J: LIM-M over two warps
set -x
O=gpurun_out
B="--no-other-configs --no-cpu-baseline"
timeout 600 python -m pytest tests/test_gpu_tail.py -m gpu -q -x > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j_pytest.log
tail -3 $O/r2j_pytest.log
timeout 300 python bench.py --workload c4 --seconds 10 $B > $O/r2j_c4_tail.json 2> $O/r2j_c4_tail.err; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name k_tail --launch-skip 20 -c 2 -o $O/r2j_tail_c4 \
  python bench.py --workload c4 --seconds 4 --steps 1 --warmup 3 $B --no-profile > $O/r2j_ncu3.log 2>&1
find $O -size +40M -delete
du -sh $O
