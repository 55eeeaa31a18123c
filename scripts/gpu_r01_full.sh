#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list and --set full capture of one timed step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
echo "== ncu launch list" 
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
echo "== ncu full (one timed step)"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -o gpurun_out/step_full -f python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/
