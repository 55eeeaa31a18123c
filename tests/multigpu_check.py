"""Multi-GPU parity check of the key all-gather + enqueue, launched with torchrun on a GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Every rank holds an identical queue replica, contributes its own keys, and after `KeyGather.enqueue` must hold exactly
what the reference's StorageQueue holds after enqueue(cat(keys_0 .. keys_{R-1})) - bit for bit, wrap-around included.
Run by tests/test_gpu_parity.py::test_multi_gpu_allgather_and_shard_parity when >= 2 GPUs are visible."""
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


def step_parity(rank, world, dev, gather):
    """DESIGN.md 7 parity definition: rank r's outputs == oracle on shard r with the shared queue snapshot; post-step
    queue == ring-buffer enqueue of the rank-ordered concatenation of every rank's keys; EMA identical on all ranks."""
    import types
    B, nf, K, D, T, m, H = 8, 2, 64, 128, 0.07, 0.999, 64
    args = types.SimpleNamespace(
        backbone=vince_b200.ResNet18, num_frames=nf, use_attention=False, feature_extractor_gpu_ids=[dev],
        pytorch_gpu_ids=[dev], vince_embedding_size=D, vince_queue_size=K, vince_temperature=T,
        vince_self_temperature=0.03, vince_momentum=m, jigsaw=False, inter_batch_comparison=True,
        self_batch_comparison=False, batch_size=B, use_imagenet=False)
    sd = vo.make_state_dict("ResNet18", D, seed=0)
    model = vince_b200.VinceModel(args)
    model.load_state_dict(sd)
    model.to(dev)
    model.train()
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(dev)
    qm.train()
    g = torch.Generator().manual_seed(100 + rank)
    data, queue_data = torch.randn((B, 3, H, H), generator=g), torch.randn((B, 3, H, H), generator=g)
    queue_init = torch.nn.functional.normalize(torch.randn((K, D), generator=torch.Generator().manual_seed(5)), dim=-1)
    queue = vince_b200.StorageQueue(K, D, device=dev)
    queue.load(queue_init.to(dev))
    queue.current_tail = K - 5                       # the gathered rows wrap around
    batch = {"data": data.to(dev), "queue_data": queue_data.to(dev), "batch_types": ["images"], "batch_sizes": [B],
             "data_source": "s", "num_frames": nf}
    kb = qm(batch, shuffle=False)[0]
    out = model.get_embeddings(batch, shuffle=False)[0]
    out.update(queue.dequeue())
    out.update({"data_source": "s", "num_frames": nf})
    out.update(kb)
    out.update(model(out))
    with torch.no_grad():
        loss = model.loss(out)["nce_loss"][1]
    qm.vince_update(model, enqueue=(queue, out["queue_embeddings"], [None] * B, "s"), gather=gather)
    torch.cuda.synchronize()
    q_sd, k_sd = vo.clone_state_dict(sd), vo.clone_state_dict(sd)
    oq = vo.StorageQueue(K, D, init=queue_init)
    ref = vo.train_step(data, queue_data, q_sd, k_sd, oq, "ResNet18", nf, T, m)
    emb_err = ((out["embeddings"].cpu() - ref["query"]["embeddings"]).norm() / ref["query"]["embeddings"].norm()).item()
    loss_err = abs(loss.item() - ref["losses"]["nce_loss"].item()) / abs(ref["losses"]["nce_loss"].item())
    ok = emb_err < 1e-3 and loss_err < 1e-3
    # queue: rank-ordered concatenation of the GPU keys of every rank, enqueued at K-5 with wrap-around
    all_keys = [torch.empty_like(out["queue_embeddings"]) for _ in range(world)]
    dist.all_gather(all_keys, out["queue_embeddings"].contiguous())
    expect = vo.StorageQueue(K, D, init=queue_init)
    expect.current_tail = K - 5
    expect.enqueue(torch.cat([k.cpu() for k in all_keys]))
    ok = ok and torch.equal(queue.vector_queue.cpu(), expect.vector_queue) and queue.current_tail == expect.current_tail
    # EMA: key-encoder parameters == oracle's (1e-6) - and therefore identical on every rank
    names = vo.vince_parameter_names(k_sd)
    got = dict(qm.queue_network.named_parameters())
    ema_err = max(((got[n].detach().cpu() - k_sd[n]).abs().max() / k_sd[n].abs().max().clamp_min(1e-12)).item() for n in names)
    ok = ok and ema_err < 1e-5
    if not ok or rank == 0:
        print("rank %d shard step: embeddings %.2e loss %.2e ema %.2e queue %s" % (
            rank, emb_err, loss_err, ema_err, "OK" if torch.equal(queue.vector_queue.cpu(), expect.vector_queue) else "MISMATCH"))
    return ok


def shuffle_bn_parity(rank, world, dev):
    """Cross-GPU shuffle-BN: with CrossGpuShuffle the batch shuffle spans the global batch, as
    nn.DataParallel's split of the shuffled batch does in the reference (vince_model.py:137-142 + :35).  Oracle: the
    global batch permuted by the same shared-seed permutation, cut into `world` slices, each slice forwarded with its own
    train-mode BatchNorm statistics, outputs un-shuffled."""
    import types
    from vince_b200.distributed import CrossGpuShuffle
    B, D, H = 8, 128, 64
    args = types.SimpleNamespace(
        backbone=vince_b200.ResNet18, num_frames=2, use_attention=False, feature_extractor_gpu_ids=[dev],
        pytorch_gpu_ids=[dev], vince_embedding_size=D, vince_queue_size=64, vince_temperature=0.07,
        vince_self_temperature=0.03, vince_momentum=0.999, jigsaw=False, inter_batch_comparison=True,
        self_batch_comparison=False, batch_size=B, use_imagenet=False)
    sd = vo.make_state_dict("ResNet18", D, seed=1)
    model = vince_b200.VinceModel(args)
    model.load_state_dict(sd)
    model.to(dev)
    model.train()
    model.cross_shuffle = CrossGpuShuffle(seed=99)
    datas = [torch.randn((B, 3, H, H), generator=torch.Generator().manual_seed(500 + r)) for r in range(world)]
    out = model.get_embeddings({"data": datas[rank].to(dev)}, shuffle=True)
    torch.cuda.synchronize()
    perm = torch.randperm(world * B, generator=torch.Generator().manual_seed(99))
    glob = torch.cat(datas)
    expect = torch.empty((world * B, D))
    with torch.no_grad():
        for d in range(world):
            sl = perm[d * B:(d + 1) * B]
            expect[sl] = vo.get_embeddings(glob[sl], vo.clone_state_dict(sd), "ResNet18", True)["embeddings"]
    mine = expect[rank * B:(rank + 1) * B]
    err = ((out["embeddings"].cpu() - mine).norm() / mine.norm()).item()
    # and it must differ from the purely local shuffle (different BatchNorm batches)
    with torch.no_grad():
        local = vo.get_embeddings(datas[rank], vo.clone_state_dict(sd), "ResNet18", True)["embeddings"]
    differs = ((local - mine).norm() / mine.norm()).item()
    if rank == 0 or err >= 1e-3:
        print("rank %d cross-GPU shuffle-BN: embeddings rel-L2 %.2e vs the global-shuffle oracle (local-BN result is %.2e away)"
              % (rank, err, differs))
    return err < 1e-3 and differs > 1e-2


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
    ok = step_parity(rank, world, dev, gather) and ok
    ok = shuffle_bn_parity(rank, world, dev) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multigpu_check world=%d: %s" % (world, "PASS" if flag.item() == 1 else "FAIL"))
    gather.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
