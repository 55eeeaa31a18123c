#!/bin/bash
# N-GPU checks (N = number of visible GPUs): bit-exact all-gather + enqueue, the bench line, the reference arm.
#   gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_multi.sh'
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== multigpu_check ($N GPUs)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py 2>&1 | tail -3 | tee gpurun_out/multigpu_check.log
echo "== bench --gpus $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 2> gpurun_out/bench_multi.err | tee gpurun_out/bench_${N}gpu.json | cut -c1-700
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_multi.err | tail -5
echo "== reference arm (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
