#!/bin/bash
# 2-GPU check of the sharded path with the final code: proofs equal the oracle (W = 2), strong-scaling bench at N = 2
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_multigpu.py -x -q -k "2" ) > gpurun_out/pytest_mgpu2.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err
tail -3 gpurun_out/pytest_mgpu2.log
