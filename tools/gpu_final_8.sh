#!/bin/bash
# round-2 final validation on EIGHT B200s of one box: the sharded north-star sweep under torchrun (strong scaling through
# the product partitioner + NCCL gather from device memory), the box-wide C-ABI handle, the 2-GPU bit-identity test
set -x
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2z_bench_8gpu.json 2> $O/r2z_bench_8gpu.err; echo "bench8 rc=$?"
timeout 300 python tools/time_multi.py --mask 0xff --steps 3 --check 32 > $O/r2z_multi_8gpu.json 2> $O/r2z_multi_8gpu.err; echo "multi8 rc=$?"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r2z_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest_multi.log
tail -3 $O/r2z_pytest_multi.log
tail -c 600 $O/r2z_bench_8gpu.err
du -sh $O
