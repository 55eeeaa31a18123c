#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/bench_4gpu.err > gpurun_out/r02_bench_line_4gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_4gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_gbs_per_rank'), 'nce', d['infonce_step_ms'], d['config']['workload'][:60], d['clocks'])" || tail -8 gpurun_out/bench_4gpu.err
echo "== pytest multi-gpu"; timeout 900 python -m pytest tests -m gpu -q -s -k "multi_gpu" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_multigpu.log
