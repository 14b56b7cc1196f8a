#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:ntt2_' -c 8 \
    -f -o gpurun_out/prof_r2e python scripts/lde_once.py 2 > gpurun_out/r2e_ncu.log 2>&1
ncu -i gpurun_out/prof_r2e.ncu-rep --page raw --csv > gpurun_out/r2e_raw.csv 2> gpurun_out/r2e_export.err
ncu -i gpurun_out/prof_r2e.ncu-rep --page source --csv -k regex:ntt2_pass1 -c 1 > gpurun_out/r2e_source_pass1.csv 2>> gpurun_out/r2e_export.err
rm -f gpurun_out/prof_r2e.ncu-rep
tail -3 gpurun_out/r2e_ncu.log; wc -c gpurun_out/r2e_*.csv
