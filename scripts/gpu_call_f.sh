#!/bin/bash
# last short call of the round: smoke(), the golden fixtures through the device path, config 4 (2^20 steps, E=16) on one GPU
mkdir -p gpurun_out
( time timeout 120 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
( time timeout 240 python -m pytest -m gpu tests/test_golden.py "tests/test_prove_gpu.py::test_large_proofs_match_c_oracle_port" -x -q ) > gpurun_out/pytest_gpu_f.log 2>&1
tail -3 gpurun_out/smoke.log; tail -4 gpurun_out/pytest_gpu_f.log
