#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'infonce_main' -c 1 -o gpurun_out/nce_main -f \
   python bench.py --config 2 --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
tail -1 gpurun_out/ncu_top.log
