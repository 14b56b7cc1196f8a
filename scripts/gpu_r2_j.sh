#!/bin/bash
# round 2, call J (2 GPUs): sharded parity on the BASELINE shapes + bench at N=2
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_multigpu.py -x -q -k "world2 or 2-" ) > gpurun_out/r2j_pytest_multigpu_w2.log 2>&1
tail -5 gpurun_out/r2j_pytest_multigpu_w2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --quick-ntt > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_n2.json'))
print(d['value'], d['e2e']['value'], d['parity_ok'], d['kernels_ms_per_step'], d['ntt'].get('sharded_lde'))
PY
