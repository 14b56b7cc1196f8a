#!/bin/bash
# 2 GPUs: bench with throughput mode; driver-style launch
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2w_bench_n2.json'))
    print('N=2', round(d['value'],4), round(d['e2e']['value'],3), d['parity_ok'], d['gpu_launches'], d['throughput_mode'], d['ntt'].get('sharded_lde'), len(d['ntt'].get('table', [])))
except Exception as e:
    print('bench failed', open('gpurun_out/r2w_bench_n2.err').read()[-1500:])
PY
