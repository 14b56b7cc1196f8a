#!/bin/bash
mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/lscpu.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b_tables.json 2> gpurun_out/bench_b_tables.err
GS_NTT_TABLES=0 timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b_notables.json 2> gpurun_out/bench_b_notables.err
tail -4 gpurun_out/pytest_gpu.log
