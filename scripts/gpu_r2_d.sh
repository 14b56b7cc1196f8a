#!/bin/bash
# round 2, call D: two-pass K1 (ntt2.cuh): parity tests, timings old vs new
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_ntt_gpu.py -x -q ) > gpurun_out/r2d_pytest_ntt.log 2>&1
tail -5 gpurun_out/r2d_pytest_ntt.log
echo "--- new (GS_NTT2=1)"; timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2d_bench_ntt_new.txt
echo "--- old (GS_NTT2=0)"; GS_NTT2=0 timeout 120 python scripts/bench_ntt.py 2>&1 | tee gpurun_out/r2d_bench_ntt_old.txt
( time timeout 600 python -m pytest tests/test_prove_gpu.py -x -q ) > gpurun_out/r2d_pytest_prove.log 2>&1
tail -3 gpurun_out/r2d_pytest_prove.log
