#!/bin/bash
# round-2 GPU pass B: GPU test tier with the fused tail, bench lines, C4 with / without the tail, ncu lists (kept small:
# gpurun_out/ over 64 MiB is not copied back at all)
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
tail -3 $O/r2b_pytest.log
timeout 900 python bench.py > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --candidates 1024 --no-other-configs > $O/r2b_bench_c5_8192.json 2> $O/r2b_bench_c5_8192.err; echo "bench8192 rc=$?"
AFSIM_TAIL=0 timeout 300 python bench.py --workload c4 --no-other-configs --no-cpu-baseline > $O/r2b_bench_c4_split.json 2> $O/r2b_bench_c4_split.err; echo "c4 split rc=$?"
timeout 300 python bench.py --workload c4 --seconds 10 --no-other-configs --no-cpu-baseline > $O/r2b_bench_c4_tail_10s.json 2> $O/r2b_bench_c4_tail_10s.err; echo "c4 tail rc=$?"
AFSIM_TAIL=2 timeout 300 python bench.py --workload c5 --no-other-configs --no-cpu-baseline --no-profile > $O/r2b_bench_c5_tail2.json 2> $O/r2b_bench_c5_tail2.err; echo "c5 tail2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2b_launches_c5_8192.csv \
  python bench.py --candidates 1024 --steps 1 --warmup 3 --no-other-configs --no-cpu-baseline --no-profile > $O/r2b_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --launch-skip 240 -c 48 -o /tmp/c5_8192_full \
  python bench.py --candidates 1024 --steps 1 --warmup 3 --no-other-configs --no-cpu-baseline --no-profile > $O/r2b_ncu2.log 2>&1
ncu -i /tmp/c5_8192_full.ncu-rep --page raw --csv > $O/r2b_c5_8192_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name k_tail --launch-skip 20 -c 2 -o $O/r2b_tail_c4 \
  python bench.py --workload c4 --seconds 4 --steps 1 --warmup 3 --no-other-configs --no-cpu-baseline --no-profile > $O/r2b_ncu3.log 2>&1
find $O -size +40M -delete
du -sh $O; ls -la $O
