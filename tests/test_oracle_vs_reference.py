"""CPU, dev container only: the oracle against the LIVE reference (imported read-only from /root/reference through
oracle/ref_loader.py + the dg_util shim) on fresh seeded inputs.  Skipped where the reference is not mounted (GPU box)."""
import pytest
import torch

import ref_loader
import vince_oracle as vo

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not mounted")


def test_similarity_cross_entropy_matches_live_reference():
    ref = ref_loader.load_reference()
    g = torch.Generator().manual_seed(3)
    for B, K, nf, T in [(12, 50, 4, 0.07), (6, 33, 2, 0.2), (5, 20, 1, 0.07)]:
        ref.loss_util.USE_FLOAT = None
        q = torch.nn.functional.normalize(torch.randn((B, 24), generator=g), dim=1)
        k = torch.nn.functional.normalize(torch.randn((B, 24), generator=g), dim=1)
        queue = torch.nn.functional.normalize(torch.randn((K, 24), generator=g), dim=1)
        fw = vo.vince_forward(q, k, queue, nf)
        r = ref.loss_util.similarity_cross_entropy(fw["vince_similarities"], T, B, 1, fw["vince_similarities_mask"])
        o = vo.similarity_cross_entropy(fw["vince_similarities"], T, fw["vince_similarities_mask"])
        assert torch.allclose(r["dists"], o["dists"], rtol=1e-6, atol=1e-6)
        assert torch.allclose(r["softmax_weight"], o["softmax_weight"], rtol=1e-6)


def test_encoder_and_queue_match_live_reference():
    ref = ref_loader.load_reference()
    args = ref_loader.make_args(backbone="ResNet18", num_frames=2, batch_size=4, queue_size=16, embedding_size=32)
    model = ref.VinceModel(args)
    sd = vo.make_state_dict("ResNet18", 32, seed=11)
    model.load_state_dict(sd, strict=True)
    model.train()
    x = torch.randn((4, 3, 40, 56), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        r = model.get_embeddings({"data": x})
    o = vo.get_embeddings(x, vo.clone_state_dict(sd), "ResNet18", True)
    assert ((o["embeddings"] - r["embeddings"]).norm() / r["embeddings"].norm()).item() < 2e-5
    q = ref.StorageQueue(10, 4, device="cpu")
    oq = vo.StorageQueue(10, 4, init=q.vector_queue.clone())
    for n in (3, 9, 10, 1):
        items = torch.randn(n, 4)
        q.enqueue(items, [None] * n, "s")
        oq.enqueue(items)
        assert torch.equal(q.vector_queue, oq.vector_queue) and q.current_tail == oq.current_tail and q.full == oq.full


def test_reference_solver_loop_runs_through_the_loader_and_matches_the_oracle_loss():
    """The harness of the GPU drop-in test (tests/test_gpu_parity.py::test_reference_solver_runs_unchanged_...), checked
    on the CPU: the reference's unmodified VinceSolver.run_train_iteration, loaded by file path through ref_loader with the
    dg_util shim, runs one training iteration over the reference's own classes, and the loss it records equals the
    oracle's restatement of the same step."""
    import collections

    from dg_util.python_utils.average_meter import RollingAverageMeter
    if not ref_loader.solver_available():
        pytest.skip("solvers/vince_solver.py not available")
    VinceSolver = ref_loader.load_reference_solver()
    ref = ref_loader.load_reference()
    B, nf, K, D, H = 4, 2, 16, 32, 32
    gen = torch.Generator().manual_seed(5)
    args = ref_loader.make_args(backbone="ResNet18", num_frames=nf, batch_size=B, queue_size=K, embedding_size=D)
    args.save_frequency = args.log_frequency = 10 ** 9
    sd = vo.make_state_dict("ResNet18", D, seed=1)
    model = ref.VinceModel(args)
    model.load_state_dict(sd, strict=True)
    model.train()
    qm = ref.VinceQueueModel(args, model)
    qm.train()
    queue = ref.StorageQueue(K, D, device="cpu")
    queue_init = queue.vector_queue.clone()
    data = torch.randn((B, 3, H, H), generator=gen)
    queue_data = torch.randn((B, 3, H, H), generator=gen)
    perm_k, perm_q = torch.randperm(B, generator=gen), torch.randperm(B, generator=gen)
    batch = {"data": data, "queue_data": queue_data, "batch_types": ["images"], "batch_sizes": [B], "num_frames": [nf],
             "data_source": ["synthetic"], "queue_data_cpu": [None] * B}
    s = object.__new__(VinceSolver)
    s.args, s.model, s.queue_model, s.vince_queue, s.use_apex = args, model, qm, queue, False
    s.optimizer = torch.optim.SGD(model.parameters(), lr=0.003, weight_decay=0.0001, momentum=0.9)
    s.time_meters = collections.defaultdict(RollingAverageMeter)
    s.loss_meters = collections.defaultdict(RollingAverageMeter)
    s.metric_meters = collections.defaultdict(RollingAverageMeter)
    s.train_logger, s.drawn_this_epoch, s.logger_iteration, s.iteration = None, True, 1, 0
    s.get_batch = lambda: (batch, None)
    real = torch.randperm
    perms = iter([perm_k, perm_q])
    torch.randperm = lambda n, *a, **k: next(perms).clone()
    try:
        s.run_train_iteration()
    finally:
        torch.randperm = real
    with torch.no_grad():
        keys = vo.get_embeddings(queue_data, vo.clone_state_dict(sd), "ResNet18", True, shuffle_order=perm_k)["embeddings"]
        q = vo.get_embeddings(data, vo.clone_state_dict(sd), "ResNet18", True, shuffle_order=perm_q)["embeddings"]
        fw = vo.vince_forward(q, keys, queue_init, nf)
        loss = vo.similarity_cross_entropy(fw["vince_similarities"], args.vince_temperature, fw["vince_similarities_mask"])["dist"]
    assert s.iteration == B and queue.current_tail == B
    assert abs(s.loss_meters["nce_loss"].history[-1] - float(loss)) < 1e-4 * abs(float(loss))
