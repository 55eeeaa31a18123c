#!/bin/bash
mkdir -p gpurun_out
echo "== ncu launch list: one training step cfg1"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_train_cfg1.csv python bench.py --config 1 --profile-train > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_train_cfg1.csv
