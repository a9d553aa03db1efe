#!/bin/bash
# round-2 validation on TWO B200s of one box: the box-wide C-ABI handle (bit identity with one GPU), the sharded
# north-star sweep under torchrun (strong scaling through the product partitioner + NCCL gather from device memory)
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r2q_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/r2q_pytest_multi.log
tail -3 $O/r2q_pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2q_bench_2gpu.json 2> $O/r2q_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 400 $O/r2q_bench_2gpu.err
timeout 200 python tools/time_multi.py --mask 0x3 --candidates 2048 --steps 2 --check 16 > $O/r2q_multi_2gpu.json 2> $O/r2q_multi_2gpu.err; echo "multi2 rc=$?"
tail -c 400 $O/r2q_multi_2gpu.err
du -sh $O
