#!/bin/bash
# final single-GPU call of the round: GPU tests, bench (both arms), ncu launch list, one ncu --set full capture
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches.csv \
    python scripts/prove_once.py 2 > gpurun_out/ncu_list.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k 'regex:ntt_pass|hash_columns|merkle_level|compose|fri_fold' -c 24 \
    -f -o gpurun_out/prof_r1d python scripts/prove_once.py 1 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_r1d.ncu-rep --page raw --csv > gpurun_out/prof_r1d_raw.csv 2> gpurun_out/ncu_export.err
sz=$(stat -c %s gpurun_out/prof_r1d.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 45000000 ]; then rm -f gpurun_out/prof_r1d.ncu-rep; echo "ncu-rep of $sz bytes dropped (pull limit); raw csv kept" >> gpurun_out/ncu_full.log; fi
tail -3 gpurun_out/pytest_gpu.log; ls -la gpurun_out | tail -12
