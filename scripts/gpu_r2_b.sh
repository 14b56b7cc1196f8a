#!/bin/bash
# round 2, call B: new bench contract (default config, quick NTT table) + "before" ncu capture of the K1 passes with warp-state sections
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --quick-ntt > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:ntt_pass' -c 12 \
    -f -o gpurun_out/prof_r2b_ntt python scripts/lde_once.py 2 > gpurun_out/r2b_ncu.log 2>&1
ncu -i gpurun_out/prof_r2b_ntt.ncu-rep --page raw --csv > gpurun_out/r2b_ntt_raw.csv 2> gpurun_out/r2b_ncu_export.err
rm -f gpurun_out/prof_r2b_ntt.ncu-rep
head -c 1200 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_ncu.log
