#!/bin/bash
# round-2: apply-epilogue / two-pass A-B: parity tests, then bench cfg2 with and without the two-pass route
mkdir -p gpurun_out
echo "== pytest (fused epilogue + full size)"; timeout 1200 python -m pytest tests -m gpu -x -q -s -k "fused_epilogue or full_size or golden or cfg4" 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|cfg4|assert" | tail -30 | tee gpurun_out/pytest_gpu_b.log
for tp in 0 1; do
echo "== bench --config 2 TWOPASS=$tp"; VINCE_B200_TWOPASS=$tp timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --profile-only 2>&1 | tail -1
done
echo "== ncu launch list cfg2 (two-pass)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_cfg2_tp.csv python bench.py --config 2 --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_cfg2_tp.csv
