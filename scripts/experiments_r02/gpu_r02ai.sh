#!/bin/bash
mkdir -p gpurun_out
echo "== compute-sanitizer memcheck: smoke + small training step + apply routes"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02_sanitizer_memcheck_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tests/train_probe.py ResNet18 8 64 2>&1 | grep -E "ERROR SUMMARY|GLOBAL|Invalid|Error" | tail -6 | tee gpurun_out/r02_sanitizer_memcheck_train.log
echo "== compute-sanitizer racecheck: smoke"
timeout 1200 compute-sanitizer --tool racecheck --print-limit 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "RACECHECK SUMMARY|smoke:|Race reported|and Read|and Write" | sort | uniq -c | sort -rn | head -20 | tee gpurun_out/r02_sanitizer_racecheck_smoke.log
