#!/bin/bash
# NOT RUN in round 2 (the GPU budget ended with the parity / A-B runs of the fused commit).  First GPU call to make next:
# the ncu launch list and a full capture of one prove with the fused commit kernels (merkle_span_kernel, leaf-hashing
# merkle_top_kernel), so that profiles/ncu_summary.json -- and with it bench.py's roofline.traffic for the class
# merkle_commit -- describes the code that is benched.  About 3 GPU-minutes.
#   gpurun --timeout 600 -- 'bash scripts/gpu_next_ncu_commit.sh'
# then, here:  python scripts/ncu_summarize.py gpurun_out/r3_full_raw.csv profiles/r3_ncu_full 26
#              and copy the per-class dram bytes into profiles/ncu_summary.json (configs.ns, with the commit hash).
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches_mimc_2e20.csv python scripts/prove_once.py 2 > gpurun_out/r3_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -c 120 -f -o gpurun_out/prof_r3 python scripts/prove_once.py 2 > gpurun_out/r3_ncu_full.log 2>&1
ncu -i gpurun_out/prof_r3.ncu-rep --page raw --csv > gpurun_out/r3_full_raw.csv 2> gpurun_out/r3_ncu_export.err
rm -f gpurun_out/prof_r3.ncu-rep
tail -2 gpurun_out/r3_ncu_list.log; tail -2 gpurun_out/r3_ncu_full.log; wc -l gpurun_out/r3_full_raw.csv gpurun_out/r3_launches_mimc_2e20.csv
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r3_pytest_gpu.log 2>&1; tail -3 gpurun_out/r3_pytest_gpu.log
