#!/bin/bash
# round-2 GPU pass E: tail with backed-off waits; ncu --set full tables of the bench shapes (raw CSV only: small)
set -x
O=gpurun_out
B="--no-other-configs --no-cpu-baseline"
timeout 600 python -m pytest tests/test_gpu_tail.py -m gpu -q -x > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
tail -3 $O/r2e_pytest.log
timeout 300 python bench.py --workload c4 --seconds 10 $B > $O/r2e_c4_tail.json 2> $O/r2e_c4_tail.err; echo "rc=$?"
timeout 300 python bench.py --workload c4 $B --no-profile > $O/r2e_c4_tail_60s.json 2> $O/r2e_c4_tail_60s.err; echo "rc=$?"
AFSIM_TAIL=2 timeout 300 python bench.py --workload c2 $B --no-profile > $O/r2e_c2_tail.json 2> $O/r2e_c2_tail.err; echo "rc=$?"
AFSIM_TAIL=2 timeout 300 python bench.py --candidates 1024 $B --no-profile > $O/r2e_c5_8192_tail.json 2> $O/r2e_c5_8192_tail.err; echo "rc=$?"
cap() {  # name, skip, count, bench args...
  name=$1; skip=$2; count=$3; shift 3
  timeout 900 ncu --set full --clock-control none --launch-skip $skip -c $count -o /tmp/$name \
    python bench.py "$@" --steps 1 --warmup 1 $B --no-profile > $O/r2e_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/r2e_${name}_raw.csv 2>/dev/null
}
cap c4 30 6 --workload c4 --seconds 4
cap c5_65536 60 26 --workload c5
cap c2 100 30 --workload c2 --seconds 10
cap c3 100 24 --workload c3 --seconds 4
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name k_tail --launch-skip 20 -c 2 -o $O/r2e_tail_c4 \
  python bench.py --workload c4 --seconds 4 --steps 1 --warmup 3 $B --no-profile > $O/r2e_ncu3.log 2>&1
find $O -size +40M -delete
du -sh $O
