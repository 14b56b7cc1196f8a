#!/bin/bash
# 8 GPUs: staged peer scatter -- parity (small, ns, 5) and bench at N = 8, 4, 2
mkdir -p gpurun_out
( time timeout 600 python scripts/shard_check.py 8 small ns 5 ) > gpurun_out/r2u_shard_check_w8.log 2>&1
grep -E "SHARD_CHECK|equals_oracle=False|FAILED|Error" gpurun_out/r2u_shard_check_w8.log | head -6
for n in 8 4 2; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2u_bench_n$n.json 2> gpurun_out/r2u_bench_n$n.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2u_bench_n$n.json'))
    print('N=$n', round(d['value'],4), round(d['e2e']['value'],3), d['parity_ok'], d['gpu_launches'], d['kernels_ms_per_step'])
except Exception as e:
    print('bench failed', open('gpurun_out/r2u_bench_n$n.err').read()[-800:])
PY
done
