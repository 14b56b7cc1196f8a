#!/bin/bash
# round 2, call A: state of the tree at round start -- GPU tests and the default bench
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest_gpu.log 2>&1
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -4 gpurun_out/r2a_pytest_gpu.log; cat gpurun_out/r2a_bench.json | head -c 1500
