#!/bin/bash
mkdir -p gpurun_out
for tp in 0 1; do
VINCE_B200_TWOPASS=$tp timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_cfg2_tp$tp.csv python bench.py --config 2 --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
done
for e in 1 2 1 2; do
echo "== bench --config 1 EPI_SETS=$e"; VINCE_B200_EPI_SETS=$e timeout 600 python bench.py --config 1 --steps 30 --warmup 5 --profile-only 2>&1 | tail -1
done
