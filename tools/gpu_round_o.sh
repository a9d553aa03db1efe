#!/bin/bash
# ring slots (chunks in flight) of the fused path at 65536 / 32768 streams per GPU
set -x
O=gpurun_out
B="--no-other-configs --no-cpu-baseline --no-profile --steps 2 --warmup 1"
for s in 4 6 8; do
  AFSIM_SLOTS=$s timeout 200 python bench.py $B > $O/r2o_c5_65536_slots$s.json 2> $O/r2o_c5_65536_slots$s.err; echo "rc=$?"
done
for s in 4 8 12; do
  AFSIM_SLOTS=$s timeout 200 python bench.py --candidates 4096 $B > $O/r2o_c5_32768_slots$s.json 2> $O/r2o_c5_32768_slots$s.err; echo "rc=$?"
done
grep -o '"ms_per_step": [0-9.]*' $O/r2o_*.json
tail -c 300 $O/r2o_c5_65536_slots8.err
