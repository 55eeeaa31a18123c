#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "== pytest multi-gpu"; timeout 900 python -m pytest tests -m gpu -q -s -k "multi_gpu" 2>&1 | tail -6 | tee gpurun_out/r02_pytest_multigpu.log
echo "== bench 2 GPUs (configs[3])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench_2gpu.err > gpurun_out/r02_bench_line_2gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line_2gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'nce', d['infonce_step_ms'], d['config']['workload'][:60])" || tail -8 gpurun_out/bench_2gpu.err
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
