#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_launches_train_cfg2.csv python bench.py --config 2 --profile-train > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
