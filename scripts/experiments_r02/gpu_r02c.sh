#!/bin/bash
mkdir -p gpurun_out
F="${1:-r50.layer3 1x1 256->1024}"
{
echo "== default"; python tests/conv_bench.py --filter "r50." --iters 7
echo "== nostats"; python tests/conv_bench.py --filter "r50." --iters 7 --nostats
echo "== statsonly"; python tests/conv_bench.py --filter "r50." --iters 7 --statsonly
echo "== skip MMA (1)"; VINCE_B200_DEBUG_SKIP_MMA=1 python tests/conv_bench.py --filter "r50." --iters 7
echo "== skip MMA+loads (7)"; VINCE_B200_DEBUG_SKIP_MMA=7 python tests/conv_bench.py --filter "r50." --iters 7
echo "== skip MMA+loads (7) nostats"; VINCE_B200_DEBUG_SKIP_MMA=7 python tests/conv_bench.py --filter "r50." --iters 7 --nostats
} 2>&1 | tee gpurun_out/conv_variants.log
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 2 -o gpurun_out/c3_full -f \
   python tests/conv_bench.py --filter "$F" --iters 3 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log; ls -la gpurun_out/*.ncu-rep
