#!/bin/bash
# round-2 check: GPU parity tests, smoke, bench lines for configs 2 / 1 / 4, ncu launch list of config 2
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|cfg4|PASS|FAIL" | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
for c in 2 1 4; do
echo "== bench --config $c"; timeout 900 python bench.py --config $c --steps 20 --warmup 5 2> gpurun_out/bench_cfg$c.err | tee gpurun_out/bench_cfg$c.json | cut -c1-600
tail -3 gpurun_out/bench_cfg$c.err
done
echo "== ncu launch list cfg2"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_cfg2.csv python bench.py --config 2 --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_cfg2.csv
