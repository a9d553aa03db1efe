#!/bin/bash
# round-2 final validation on ONE B200: GPU test tier, smoke, the default bench line, launch list of the default line
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest.log
tail -4 $O/r2z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2z_smoke.log
timeout 900 python bench.py > $O/r2z_bench.json 2> $O/r2z_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2z_bench_reference.json 2> $O/r2z_bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --candidates 1024 --no-other-configs > $O/r2z_bench_c5_8192.json 2> $O/r2z_bench_c5_8192.err; echo "bench8192 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-other-configs --no-cpu-baseline --no-profile > $O/r2z_ncu1.log 2>&1
find $O -size +40M -delete
du -sh $O
