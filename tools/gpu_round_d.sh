#!/bin/bash
# round-2 GPU pass D: tail kernel with the small code footprint; chunk-size experiments on the per-rank C5 shape
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tail.py tests/test_gpu_parity.py -m gpu -q -x > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
tail -3 $O/r2d_pytest.log
B="--no-other-configs --no-cpu-baseline"
timeout 300 python bench.py --workload c4 --seconds 10 $B > $O/r2d_c4_tail.json 2> $O/r2d_c4_tail.err; echo "rc=$?"
AFSIM_TAIL=2 timeout 300 python bench.py --workload c2 $B --no-profile > $O/r2d_c2_tail.json 2> $O/r2d_c2_tail.err; echo "rc=$?"
AFSIM_TAIL=2 timeout 300 python bench.py --candidates 1024 $B --no-profile > $O/r2d_c5_8192_tail.json 2> $O/r2d_c5_8192_tail.err; echo "rc=$?"
for c in 480 960 2880; do
  AFSIM_CHUNK=$c timeout 300 python bench.py --candidates 1024 $B --no-profile > $O/r2d_c5_8192_chunk$c.json 2> $O/r2d_c5_8192_chunk$c.err; echo "rc=$?"
done
AFSIM_CHUNK=480 AFSIM_SLOTS=52 timeout 300 python bench.py --candidates 1024 $B --no-profile > $O/r2d_c5_8192_chunk480_s52.json 2> /dev/null; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name k_tail --launch-skip 20 -c 2 -o $O/r2d_tail_c4 \
  python bench.py --workload c4 --seconds 4 --steps 1 --warmup 3 $B --no-profile > $O/r2d_ncu3.log 2>&1
find $O -size +40M -delete
du -sh $O
