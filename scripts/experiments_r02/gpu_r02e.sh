#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== pytest ($N GPUs)"; timeout 1200 python -m pytest tests -m gpu -q -s -k "multi_gpu or knn or fused_epilogue or queue_shadow or loss_backward or jigsaw" 2>&1 | grep -E "rel-L2|oracle|passed|failed|rror|kNN|assert|FAILED|multigpu|shuffle|rank" | tail -30 | tee gpurun_out/pytest_gpu_e.log
if [ "$N" -gt 1 ]; then
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_multi.err | tee gpurun_out/bench_${N}gpu.json | cut -c1-900
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_multi.err | tail -5
fi
