#!/bin/bash
mkdir -p gpurun_out
echo "--- K1 tests with GS_NTT2_SHFL=1"; ( GS_NTT2_SHFL=1 timeout 300 python -m pytest tests/test_ntt_gpu.py -x -q ) 2>&1 | tail -2
echo "--- bench_ntt plain loads, smem exchange (GS_NTT2_TMA=0)"; GS_NTT2_TMA=0 timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2_shfl_off.txt | sed -n 2,4p
echo "--- bench_ntt plain loads, shuffle exchange (GS_NTT2_TMA=0 GS_NTT2_SHFL=1)"; GS_NTT2_TMA=0 GS_NTT2_SHFL=1 timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2_shfl_on.txt | sed -n 2,4p
GS_NTT2_SHFL=1 timeout 200 ncu --set full --clock-control none -k 'regex:ntt2_pass2' -c 2 -f -o gpurun_out/prof_shfl python scripts/lde_once.py 1 > gpurun_out/r2_shfl_ncu.log 2>&1
ncu -i gpurun_out/prof_shfl.ncu-rep --page raw --csv > gpurun_out/r2_shfl_raw.csv 2> /dev/null
rm -f gpurun_out/prof_shfl.ncu-rep
