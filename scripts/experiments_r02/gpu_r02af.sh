#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-train 2> gpurun_out/bench_8gpu.err > gpurun_out/r02_bench_line_8gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_8gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_gbs_per_rank'), 'nce', d['infonce_step_ms'], d['config']['workload'][:60], d['clocks'])" || tail -8 gpurun_out/bench_8gpu.err
