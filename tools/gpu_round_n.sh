#!/bin/bash
# per-rank shapes of the strong-scaled north-star sweep at N = 4 (16384 streams per GPU): split kernels (default) against fused
set -x
O=gpurun_out
B="--no-other-configs --no-cpu-baseline --no-profile"
timeout 200 python bench.py --candidates 2048 $B > $O/r2n_c5_16384_default.json 2> $O/r2n_c5_16384_default.err; echo "rc=$?"
AFSIM_SPLIT=1 timeout 200 python bench.py --candidates 2048 $B > $O/r2n_c5_16384_fused.json 2> $O/r2n_c5_16384_fused.err; echo "rc=$?"
AFSIM_SPLIT=2 timeout 200 python bench.py --candidates 2048 $B > $O/r2n_c5_16384_split.json 2> $O/r2n_c5_16384_split.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*' $O/r2n_c5_16384_*.json
