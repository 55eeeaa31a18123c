#!/bin/bash
mkdir -p gpurun_out
{
echo "== stem default"; timeout 120 python tests/elem_bench.py --only stem | grep conv
echo "== stem SKIP_MMA=1 (loads+epilogue)"; VINCE_B200_DEBUG_SKIP_MMA=1 timeout 120 python tests/elem_bench.py --only stem | grep conv
echo "== stem SKIP_MMA=7 (epilogue only)"; VINCE_B200_DEBUG_SKIP_MMA=7 timeout 120 python tests/elem_bench.py --only stem | grep conv
echo "== stem SKIP_MMA=6 (MMA+epilogue, no loads)"; VINCE_B200_DEBUG_SKIP_MMA=6 timeout 120 python tests/elem_bench.py --only stem | grep conv
echo "== stem RESIDENT=0"; VINCE_B200_RESIDENT=0 timeout 120 python tests/elem_bench.py --only stem | grep conv
echo "== stem CTA_PAIR=0"; VINCE_B200_CTA_PAIR=0 timeout 120 python tests/elem_bench.py --only stem | grep conv
} 2>&1 | tee gpurun_out/exp.log
