#!/bin/bash
mkdir -p gpurun_out
echo "--- tests with GS_NTT2_TMA=1"; ( GS_NTT2_TMA=1 timeout 300 python -m pytest tests/test_ntt_gpu.py -x -q ) 2>&1 | tail -3
echo "--- plain loads"; timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2h_bench_ntt_plain.txt
echo "--- GS_NTT2_TMA=1 (pass 2 bulk prefetch)"; GS_NTT2_TMA=1 timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2h_bench_ntt_tma1.txt
GS_NTT2_TMA=1 timeout 300 ncu --set full --clock-control none -k 'regex:ntt2_' -c 4 -f -o gpurun_out/prof_r2h python scripts/lde_once.py 2 > gpurun_out/r2h_ncu.log 2>&1
ncu -i gpurun_out/prof_r2h.ncu-rep --page raw --csv > gpurun_out/r2h_raw.csv 2> /dev/null
rm -f gpurun_out/prof_r2h.ncu-rep
