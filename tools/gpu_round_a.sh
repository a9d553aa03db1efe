#!/bin/bash
# round-2 GPU pass A: GPU test tier, default bench line (C5 65536 pairs + sub-records), the per-rank C5 shape, ncu lists
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"
python bench.py --candidates 1024 --no-other-configs > $O/r2a_bench_c5_8192.json 2> $O/r2a_bench_c5_8192.err; echo "bench8192 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2a_launches_c5_8192.csv \
  python bench.py --candidates 1024 --steps 1 --warmup 3 --no-other-configs --no-cpu-baseline --no-profile > $O/r2a_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 250 -c 120 -o $O/r2a_c5_8192_full \
  python bench.py --candidates 1024 --steps 1 --warmup 3 --no-other-configs --no-cpu-baseline --no-profile > $O/r2a_ncu2.log 2>&1
ls -la $O
