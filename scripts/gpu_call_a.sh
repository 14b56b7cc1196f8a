#!/bin/bash
# one GPU-box call: pipe probes, full GPU test-suite, bench with and without the full twiddle tables
mkdir -p gpurun_out
scripts/pipe_probe.bin > gpurun_out/pipe_probe.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tables.json 2> gpurun_out/bench_tables.err
GS_NTT_TABLES=0 timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_notables.json 2> gpurun_out/bench_notables.err
tail -4 gpurun_out/pytest_gpu.log
