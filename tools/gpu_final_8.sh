#!/bin/bash
# round-2 final validation on EIGHT B200s of one box: the sharded north-star sweep under torchrun (strong scaling through
# the product partitioner + NCCL gather from device memory) and the box-wide C-ABI handle (one process, no Python sharding)
set -x
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --steps 3 --warmup 3 > $O/r2z_bench_8gpu.json 2> $O/r2z_bench_8gpu.err; echo "bench8 rc=$?"
timeout 200 python tools/time_multi.py --mask 0xff --steps 2 --check 16 > $O/r2z_multi_8gpu.json 2> $O/r2z_multi_8gpu.err; echo "multi8 rc=$?"
tail -c 400 $O/r2z_bench_8gpu.err; tail -c 300 $O/r2z_multi_8gpu.err
du -sh $O
