#!/bin/bash
mkdir -p gpurun_out
echo "== bench cfg1 no-train"; timeout 900 python bench.py --config 1 --steps 10 --warmup 3 --no-train 2> gpurun_out/bench_g1.err > gpurun_out/bench_g1.json; python -c "
import json
d=json.load(open('gpurun_out/bench_g1.json')); print(d['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['roofline']['ms_per_step'], d['roofline']['avg_launch_us'], d['clocks'], d['e2e']['ms_per_step'], d['gpu_launches_per_step'])"
echo "== ncu launch list cfg1"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_cfg1_g.csv python bench.py --config 1 --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_cfg1_g.csv
