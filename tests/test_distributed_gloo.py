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


def test_cross_gpu_shuffle_plan_is_a_consistent_global_permutation():
    """CrossGpuShuffle.plan on CPU: simulate the all-to-all for every rank pair and check that rank r ends up with frames
    perm[r*B:(r+1)*B] in order, and that the reverse exchange returns every output to its owner and position."""
    from vince_b200.distributed import CrossGpuShuffle
    for world, B, seed in ((2, 8, 0), (4, 6, 1), (8, 4, 2), (3, 5, 3)):
        perm = torch.randperm(world * B, generator=torch.Generator().manual_seed(seed))
        plans = [CrossGpuShuffle.plan(perm, world, r, B) for r in range(world)]
        data = [torch.arange(r * B, (r + 1) * B) for r in range(world)]            # frame ids
        # forward all-to-all: rank s sends data[s][send_order] split by send_counts; rank d receives grouped by source
        shuffled = []
        for d in range(world):
            recv = []
            for s in range(world):
                so, sc, rc, _ = plans[s]
                off = sum(sc[:d])
                chunk = data[s][so][off:off + sc[d]]
                assert len(chunk) == plans[d][2][s]                                 # counts agree on both sides
                recv.append(chunk)
            got = torch.cat(recv)
            out = torch.empty_like(got)
            out[plans[d][3]] = got
            shuffled.append(out)
            assert torch.equal(out, perm[d * B:(d + 1) * B])
        # reverse: outputs y = 10 * frame id travel back
        for s in range(world):
            so, sc, rc, place = plans[s]
            back = []
            for d in range(world):
                y = (10 * shuffled[d])[plans[d][3]]                                 # grouped by source again
                off = sum(plans[d][2][:s])
                back.append(y[off:off + plans[d][2][s]])
            back = torch.cat(back)
            out = torch.empty_like(back)
            out[so] = back
            assert torch.equal(out, 10 * data[s])


def _shuffle_worker(rank, world, port, B, ret):
    sys.path.insert(0, ROOT)
    from vince_b200.distributed import CrossGpuShuffle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cs = CrossGpuShuffle(seed=7)
    ok = True
    for step in range(3):
        x = torch.arange(rank * B, (rank + 1) * B, dtype=torch.float32).view(B, 1).repeat(1, 3) + 100 * step
        sh, ctx = cs.exchange(x)
        perm = ctx[4]
        ok = ok and torch.equal(sh[:, 0] - 100 * step, perm[rank * B:(rank + 1) * B].float())
        y = cs.restore(sh * 2.0, ctx)
        ok = ok and torch.equal(y, x * 2.0)
    ret[rank] = ok
    dist.destroy_process_group()


def test_cross_gpu_shuffle_exchange_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 1000) + 77
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_shuffle_worker, args=(world, port, 6, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
