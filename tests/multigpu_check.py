"""Multi-GPU parity check of the key all-gather + enqueue, launched with torchrun on a GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Every rank holds an identical queue replica, contributes its own keys, and after `KeyGather.enqueue` must hold exactly
what the reference's StorageQueue holds after enqueue(cat(keys_0 .. keys_{R-1})) - bit for bit, wrap-around included.
Not collected by pytest (needs >= 2 GPUs)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vince_oracle as vo  # noqa: E402

import vince_b200  # noqa: E402
from vince_b200.distributed import KeyGather  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = "cuda:%d" % local
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(dev))
    gather = KeyGather(dev)
    ok = True
    for K, D, n_local in [(1024, 128, 64), (100, 32, 24), (4096, 128, 256)]:
        init = torch.nn.functional.normalize(torch.randn((K, D), generator=torch.Generator().manual_seed(7)), dim=-1)
        queue = vince_b200.StorageQueue(K, D, device=dev)
        queue.load(init.to(dev))
        oracle = vo.StorageQueue(K, D, init=init)
        for step in range(2 * (K // (n_local * world)) + 3):
            keys = torch.randn((n_local, D), generator=torch.Generator().manual_seed(1000 * step + rank)).to(dev)
            gather.enqueue(queue, keys, None, "src")
            oracle.enqueue(torch.cat([torch.randn((n_local, D), generator=torch.Generator().manual_seed(1000 * step + r))
                                      for r in range(world)]))
            torch.cuda.synchronize()
            same = torch.equal(queue.vector_queue.cpu(), oracle.vector_queue)
            state = queue.current_tail == oracle.current_tail and queue.full == oracle.full
            if not (same and state):
                ok = False
                print("rank %d K=%d step %d MISMATCH same=%s tail=%d/%d" % (rank, K, step, same, queue.current_tail, oracle.current_tail))
                break
        shadow = torch.empty_like(queue.vector_queue)
        vince_b200.ops.round_tf32(queue.vector_queue, shadow)
        ok = ok and torch.equal(shadow, queue.vector_queue_tf32)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multigpu_check world=%d: %s" % (world, "PASS" if flag.item() == 1 else "FAIL"))
    gather.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
