#!/bin/bash
# round-2 GPU pass C: full GPU test tier, C4 with the restructured tail vs the split kernels, C2 / C5 with the tail forced
set -x
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
tail -5 $O/r2c_pytest.log
B="--no-other-configs --no-cpu-baseline"
timeout 300 python bench.py --workload c4 --seconds 10 $B > $O/r2c_c4_tail.json 2> $O/r2c_c4_tail.err; echo "rc=$?"
AFSIM_TAIL=1 timeout 300 python bench.py --workload c4 --seconds 10 $B > $O/r2c_c4_split.json 2> $O/r2c_c4_split.err; echo "rc=$?"
timeout 300 python bench.py --workload c2 $B > $O/r2c_c2_split.json 2> $O/r2c_c2_split.err; echo "rc=$?"
AFSIM_TAIL=2 timeout 300 python bench.py --workload c2 $B > $O/r2c_c2_tail.json 2> $O/r2c_c2_tail.err; echo "rc=$?"
timeout 300 python bench.py --candidates 1024 $B > $O/r2c_c5_8192_split.json 2> $O/r2c_c5_8192_split.err; echo "rc=$?"
AFSIM_TAIL=2 timeout 300 python bench.py --candidates 1024 $B > $O/r2c_c5_8192_tail.json 2> $O/r2c_c5_8192_tail.err; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name k_tail --launch-skip 20 -c 2 -o $O/r2c_tail_c4 \
  python bench.py --workload c4 --seconds 4 --steps 1 --warmup 3 $B --no-profile > $O/r2c_ncu3.log 2>&1
find $O -size +40M -delete
du -sh $O
