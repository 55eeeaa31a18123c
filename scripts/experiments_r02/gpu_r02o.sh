#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/apply1_l1 -f \
   python tests/conv_bench.py --filter "r50.layer1 1x1 64->256" --apply 1 --iters 2 > gpurun_out/ncu_apply.log 2>&1
tail -2 gpurun_out/ncu_apply.log
