#!/bin/bash
# 8 GPUs: BASELINE configs 4 and 5 sharded over 8 ranks (proof bytes against the C oracle), then the bench at N = 8 and N = 4
mkdir -p gpurun_out
( time timeout 900 python scripts/shard_check.py 8 small 4 5 ns ) > gpurun_out/r2t_shard_check_w8.log 2>&1
grep -E "SHARD_CHECK|C oracle proof|equals_oracle=False|FAILED|Error" gpurun_out/r2t_shard_check_w8.log | head -12; tail -2 gpurun_out/r2t_shard_check_w8.log
for n in 8 4; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2t_bench_n$n.json 2> gpurun_out/r2t_bench_n$n.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2t_bench_n$n.json'))
    print('N=$n', round(d['value'],4), round(d['e2e']['value'],3), d['parity_ok'], d['gpu_launches'], d['kernels_ms_per_step'], d['ntt'].get('sharded_lde'))
except Exception as e:
    print('bench failed', open('gpurun_out/r2t_bench_n$n.err').read()[-800:])
PY
done
