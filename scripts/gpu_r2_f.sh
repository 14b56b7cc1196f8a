#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_ntt_gpu.py -x -q ) 2>&1 | tail -2
timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2f_bench_ntt.txt
timeout 300 ncu --set full --clock-control none -k 'regex:ntt2_' -c 4 -f -o gpurun_out/prof_r2f python scripts/lde_once.py 2 > gpurun_out/r2f_ncu.log 2>&1
ncu -i gpurun_out/prof_r2f.ncu-rep --page raw --csv > gpurun_out/r2f_raw.csv 2> /dev/null
rm -f gpurun_out/prof_r2f.ncu-rep
