#!/bin/bash
# round 2, final single-GPU call: GPU tests, bench (both arms, every config), ncu launch list and one full capture of a prove
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
for cfg in 2 3 4 5; do
  timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --quick-ntt > gpurun_out/r2_bench_cfg$cfg.json 2> gpurun_out/r2_bench_cfg$cfg.err
done
python - <<'PY'
import json
for name in ('n1', 'cfg2', 'cfg3', 'cfg4', 'cfg5'):
    try:
        d = json.load(open(f'gpurun_out/r2_bench_{name}.json'))
        print(name, 'value', round(d['value'], 4), 'e2e', round(d['e2e']['value'], 3), 'parity', d['parity_ok'], 'launches', d['gpu_launches'], 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'], 4),
              'cpu', round(d['cpu_baseline']['value'], 1), 'trace', d['e2e']['stages_ms'][0])
    except Exception as e:
        print(name, 'FAILED', e)
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_mimc_2e20.csv python scripts/prove_once.py 2 > gpurun_out/r2_ncu_list.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -c 160 -f -o gpurun_out/prof_r2_final python scripts/prove_once.py 2 > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/prof_r2_final.ncu-rep --page raw --csv > gpurun_out/r2_final_raw.csv 2> gpurun_out/r2_ncu_export.err
rm -f gpurun_out/prof_r2_final.ncu-rep
tail -2 gpurun_out/r2_ncu_list.log; tail -2 gpurun_out/r2_ncu_full.log; wc -l gpurun_out/r2_final_raw.csv gpurun_out/r2_launches_mimc_2e20.csv
