#!/bin/bash
# 2 GPUs: peer barrier + root scatter; gather threshold sweep
mkdir -p gpurun_out
( time timeout 600 python scripts/shard_check.py 2 small ns 4 ) > gpurun_out/r2s_shard_check_w2.log 2>&1
grep -E "SHARD_CHECK|equals_oracle=False|FAILED|Error" gpurun_out/r2s_shard_check_w2.log | head; tail -3 gpurun_out/r2s_shard_check_w2.log
for g in 21 19; do
GS_SHARD_GATHER_LOG=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$((g-19)) bench.py --gpus 2 --steps 5 --warmup 3 --quick-ntt > gpurun_out/r2s_bench_n2_g$g.json 2> gpurun_out/r2s_bench_n2_g$g.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2s_bench_n2_g$g.json'))
    print('gather_log=$g', round(d['value'],4), round(d['e2e']['value'],3), d['parity_ok'], d['gpu_launches'], d['kernels_ms_per_step'])
except Exception as e:
    print('bench failed', open('gpurun_out/r2s_bench_n2_g$g.err').read()[-800:])
PY
done
