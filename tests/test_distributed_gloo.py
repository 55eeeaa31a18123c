"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU key exchange.  Every rank
all-gathers its keys, places the rank-ordered rows with `ring_slices` (the arithmetic KeyGather / the C function
vince_allgather_enqueue use) and must end with the queue the reference's StorageQueue would hold after
enqueue(cat(keys_0, keys_1)) - including wrap-around.  The data-path collective itself (NCCL) needs GPUs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, K, D, n_local, steps, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vince_oracle as vo
    from vince_b200.distributed import ring_slices
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    init = torch.randn((K, D), generator=g)
    queue = init.clone()
    tail, full = 0, False
    oracle = vo.StorageQueue(K, D, init=init)
    ok = True
    for step in range(steps):
        gk = torch.Generator().manual_seed(100 * step + rank)
        keys = torch.randn((n_local, D), generator=gk)
        gathered = [torch.empty_like(keys) for _ in range(world)]
        dist.all_gather(gathered, keys)
        allk = torch.cat(gathered)                               # rank order
        slices, tail, wrapped = ring_slices(tail, allk.shape[0], K)
        full = full or wrapped
        for src, dst, cnt in slices:
            queue[dst:dst + cnt] = allk[src:src + cnt]
        expect = torch.cat([torch.randn((n_local, D), generator=torch.Generator().manual_seed(100 * step + r))
                            for r in range(world)])
        oracle.enqueue(expect)
        ok = ok and torch.equal(queue, oracle.vector_queue) and tail == oracle.current_tail and full == oracle.full
    # replicas identical across ranks
    other = [torch.empty_like(queue) for _ in range(world)]
    dist.all_gather(other, queue)
    ok = ok and all(torch.equal(o, queue) for o in other)
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("K,n_local", [(64, 8), (40, 12), (32, 16)])
def test_allgather_enqueue_host_logic_world2(K, n_local):
    world = 2
    port = 29500 + (os.getpid() % 1000) + K
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, K, 8, n_local, 7, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
