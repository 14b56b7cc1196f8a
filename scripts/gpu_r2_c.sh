#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/ntt_lab.bin > gpurun_out/r2c_ntt_lab.txt 2>&1
cat gpurun_out/r2c_ntt_lab.txt
