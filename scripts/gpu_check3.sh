#!/bin/bash
mkdir -p gpurun_out
echo "== probes (resident weights forced)"; VINCE_B200_RESIDENT=2 timeout 600 python tests/gpu_probe.py gemm conv stem 2>&1 | grep -v PASS | tail -30 | tee gpurun_out/probe_res.log
echo "== probes"; timeout 600 python tests/gpu_probe.py gemm conv stem 2>&1 | grep -v PASS | tail -30 | tee gpurun_out/probe.log
echo "== conv bench"; timeout 300 python tests/conv_bench.py --filter layer1 2>&1 | tail -20 | tee gpurun_out/conv_bench.log
VINCE_B200_RESIDENT=0 timeout 300 python tests/conv_bench.py --filter layer1 2>&1 | tail -20 | tee -a gpurun_out/conv_bench.log
echo "== elem bench"; timeout 300 python tests/elem_bench.py --only stem 2>&1 | tail -20 | tee gpurun_out/elem_bench.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
