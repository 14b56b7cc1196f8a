#!/bin/bash
# 2-GPU experiment: NCCL point-to-point channel count for the digest all-to-all
mkdir -p gpurun_out
for ch in 16 32; do
  NCCL_MIN_P2P_NCHANNELS=$ch NCCL_MAX_P2P_NCHANNELS=32 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((ch/16)) \
      bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_p2pch$ch.json 2> gpurun_out/bench_n2_p2pch$ch.err
done
