#!/bin/bash
# 2 GPUs: sharded parity on the BASELINE shapes (gather-and-replicate FRI, graphs on the sharded path) + bench at N=2
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests/test_multigpu.py -x -q ) > gpurun_out/r2q_pytest_multigpu_w2.log 2>&1
tail -6 gpurun_out/r2q_pytest_multigpu_w2.log
for g in 1 0; do
GS_SHARD_GRAPHS=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus 2 --steps 5 --warmup 3 --quick-ntt > gpurun_out/r2q_bench_n2_g$g.json 2> gpurun_out/r2q_bench_n2_g$g.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2q_bench_n2_g$g.json'))
    print('graphs=$g', round(d['value'],4), round(d['e2e']['value'],3), d['parity_ok'], d['gpu_launches'], d['kernels_ms_per_step'])
except Exception as e:
    print('bench failed', open('gpurun_out/r2q_bench_n2_g$g.err').read()[-600:])
PY
done
