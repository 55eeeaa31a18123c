#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/l1_default -f python tests/conv_bench.py --filter r18.layer1 --iters 2 > gpurun_out/ncu_l1.log 2>&1
VINCE_B200_DEBUG_SKIP_MMA=7 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/l1_skel -f python tests/conv_bench.py --filter r18.layer1 --iters 2 >> gpurun_out/ncu_l1.log 2>&1
tail -3 gpurun_out/ncu_l1.log; ls -la gpurun_out/*.ncu-rep
