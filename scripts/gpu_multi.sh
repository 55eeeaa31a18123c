#!/bin/bash
mkdir -p gpurun_out
echo "== multigpu_check"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py 2>&1 | tail -8 | tee gpurun_out/multigpu_check.log
echo "== bench --gpus 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 2> gpurun_out/bench2.err | tee gpurun_out/bench_2gpu.json
tail -5 gpurun_out/bench2.err
echo "== reference arm (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -2
