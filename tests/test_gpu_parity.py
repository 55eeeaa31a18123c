"""Parity tests proper: the CUDA path (through the reference-shaped Python API -> ctypes -> C ABI) against
  (1) the golden vectors generated from the unmodified reference (tests/golden, oracle/make_golden.py),
  (2) the CPU oracle (oracle/vince_oracle.py) on seeded inputs,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances: 1e-3 relative on fp32 embeddings / loss (BASELINE.json north_star); bit-exact for the ring buffer;
1e-6 relative for the EMA.  Nothing here reads /root/reference.
"""
import contextlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vince_oracle as vo
from conftest import make_args

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EMB_TOL = 1e-3      # rel-L2 on embeddings (north_star)
LOSS_TOL = 1e-3     # relative on the loss (north_star)


def rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a)).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def checksum(t):
    t = t.detach().double().flatten().cpu()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder(97.0) + 1.0
    return np.array([t.sum().item(), (t * w).sum().item(), t.abs().sum().item()])


def sd_checksum(sd):
    acc = np.zeros(3)
    for v in sd.values():
        if v.is_floating_point():
            acc += checksum(v)
    return acc


@contextlib.contextmanager
def injected_randperm(perms):
    real = torch.randperm
    it = iter(perms)

    def fake(n, *a, **k):
        p = next(it)
        assert p.numel() == n
        dev = k.get("device", None)
        return p.clone().to(dev) if dev is not None else p.clone()

    torch.randperm = fake
    try:
        yield
    finally:
        torch.randperm = real


def build_model(backbone, nf, B, K, D, T=0.07, seed=0, jigsaw=False, ibc=True, self_cmp=False, passes=3):
    import vince_b200
    args = make_args(backbone=backbone, num_frames=nf, batch_size=B, queue_size=K, embedding_size=D, temperature=T,
                     jigsaw=jigsaw, inter_batch_comparison=ibc, self_batch_comparison=self_cmp, device=DEV, passes=passes)
    model = vince_b200.VinceModel(args)
    sd = vo.make_state_dict(backbone, D, jigsaw=jigsaw, seed=seed)
    model.load_state_dict(sd, strict=True)
    model.to(DEV)
    model.train()
    return args, model, sd


# ----------------------------------------------------------------------------------------------------------
# encoder vs golden vectors from the reference
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,backbone", [("r18_64", "ResNet18"), ("r50_64", "ResNet50"), ("r18_odd", "ResNet18")])
def test_encoder_matches_reference_golden(golden, case, backbone):
    g = golden("encoder.npz")
    B, nf, H, W, D, seed, shuffle = [int(v) for v in g[case + "/cfg"]]
    args, model, sd = build_model(backbone, nf, B, 64, D, seed=seed)
    np.testing.assert_allclose(sd_checksum(sd), g[case + "/weights_checksum"], rtol=1e-9)
    x = torch.from_numpy(g[case + "/x"]).to(DEV)
    inputs = {"data": x, "batch_types": ["images"], "batch_sizes": [B]}
    if shuffle:
        with injected_randperm([torch.from_numpy(g[case + "/perm"])]):
            r = model.get_embeddings(inputs, shuffle=True)[0]
    else:
        r = model.get_embeddings(inputs, shuffle=False)[0]
    torch.cuda.synchronize()
    assert rel(r["embeddings"], g[case + "/embeddings"]) < EMB_TOL
    assert rel(r["prenorm_features"], g[case + "/prenorm_features"]) < EMB_TOL
    # intermediate features are looser than the north-star quantities (embeddings / loss): these golden cases are
    # deliberately ill-conditioned (B=4 at 64x64 leaves 16 samples per channel for layer4's batch statistics, and a
    # random-init ResNet-50 amplifies rounding ~20x) and the reference's own fp32 result is itself
    # ~1e-4 away from an fp64 evaluation of the same network.
    feat_tol = 3 * EMB_TOL if backbone == "ResNet50" else EMB_TOL
    assert rel(r["extracted_features"], g[case + "/extracted_features"]) < feat_tol
    assert rel(r["spatial_features"], g[case + "/spatial_features"]) < feat_tol
    assert r["spatial_features"].shape == g[case + "/spatial_features"].shape
    post = model.state_dict()
    assert rel(post["feature_extractor.module.model.bn1.running_mean"], g[case + "/bn1_running_mean"]) < 1e-4
    assert rel(post["feature_extractor.module.model.bn1.running_var"], g[case + "/bn1_running_var"]) < 1e-4
    last_rv = [k for k in post if k.endswith("running_var")][-1]
    assert rel(post[last_rv], g[case + "/last_bn_running_var"]) < 1e-3
    assert int(post["feature_extractor.module.model.bn1.num_batches_tracked"]) == int(g[case + "/bn1_num_batches"])
    # eval-mode BN with the updated running statistics
    model.eval()
    r_eval = model.get_embeddings({"data": x})
    torch.cuda.synchronize()
    assert rel(r_eval["embeddings"], g[case + "/eval_embeddings"]) < EMB_TOL
    assert rel(r_eval["extracted_features"], g[case + "/eval_extracted_features"]) < feat_tol


def test_ema_matches_reference_golden(golden):
    import vince_b200
    g = golden("encoder.npz")
    case = "r18_64"
    B, nf, H, W, D, seed, shuffle = [int(v) for v in g[case + "/cfg"]]
    args, model, sd = build_model("ResNet18", nf, B, 64, D, seed=seed)
    # the golden EMA was taken after one train-mode forward (BN stats are not EMA'd, so only weights matter)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    with torch.no_grad():
        for i, p in enumerate(model.vince_parameters()):
            p.add_(0.01 * ((i % 7) - 3))
    qm.param_update(model, 0.999)
    torch.cuda.synchronize()
    key_params = qm.queue_network.vince_parameters()
    n_tensors, n_elems = [int(v) for v in g[case + "/ema_n_tensors"]]
    assert len(key_params) == n_tensors and sum(p.numel() for p in key_params) == n_elems
    np.testing.assert_allclose(sum(checksum(p) for p in key_params), g[case + "/ema_checksum"], rtol=1e-6)
    got = dict(qm.queue_network.named_parameters())["embedding.2.bias"]
    assert rel(got, g[case + "/ema_embedding2_bias"]) < 1e-6
    # momentum 0 == hard copy (vince_solver.py:296,318)
    qm.param_update(model, 0.0)
    torch.cuda.synchronize()
    for a, b in zip(qm.queue_network.vince_parameters(), model.vince_parameters()):
        assert torch.equal(a, b)


def test_jigsaw_matches_reference_golden(golden):
    g = golden("jigsaw.npz")
    case = "r18_jigsaw"
    B, nf, H, W, D, seed = [int(v) for v in g[case + "/cfg"]]
    args, model, sd = build_model("ResNet18", nf, B, 64, D, seed=seed, jigsaw=True)
    np.testing.assert_allclose(sd_checksum(sd), g[case + "/weights_checksum"], rtol=1e-9)
    x = torch.from_numpy(g[case + "/x"]).to(DEV)
    perm = torch.from_numpy(g[case + "/perm"])
    orders = torch.from_numpy(g[case + "/orders"])
    # the reference draws randperm(9) per row on the host (vince_model.py:166); here the per-row orders are one batched
    # device draw, so the golden orders are handed in explicitly
    with injected_randperm([perm]):
        r = model.get_embeddings({"data": x, "batch_types": ["images"], "batch_sizes": [B]}, jigsaw=True, shuffle=True,
                                 jigsaw_orders=orders)[0]
    torch.cuda.synchronize()
    assert rel(r["embeddings"], g[case + "/embeddings"]) < EMB_TOL
    assert rel(r["prenorm_features"], g[case + "/prenorm_features"]) < EMB_TOL


# ----------------------------------------------------------------------------------------------------------
# InfoNCE vs golden vectors from the reference
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["ibc_nf4", "ibc_nf2", "ibc_nf1", "ibc_self_nf4", "moco", "ibc_short_batch"])
def test_infonce_matches_reference_golden(golden, case):
    import vince_b200
    g = golden("infonce.npz")
    B, K, D, nf, ibc, self_cmp, Bact = [int(v) for v in g[case + "/cfg"]]
    T, T_self = [float(v) for v in g[case + "/T"]]
    args = make_args(num_frames=nf, batch_size=B, queue_size=K, embedding_size=D, temperature=T, self_temperature=T_self,
                     inter_batch_comparison=bool(ibc), self_batch_comparison=bool(self_cmp), device=DEV)
    model = vince_b200.VinceModel.__new__(vince_b200.VinceModel)          # loss path never touches the network
    torch.nn.Module.__init__(model)
    model.args, model.num_frames, model.launches, model._device = args, nf, 0, DEV
    q = torch.from_numpy(g[case + "/q"]).to(DEV)
    k = torch.from_numpy(g[case + "/k"]).to(DEV)
    queue = torch.from_numpy(g[case + "/queue"]).to(DEV)
    inputs = {"extracted_features": q, "embeddings": q, "queue_embeddings": k, "queue_vectors": queue,
              "data_source": "synthetic", "num_frames": nf}
    out = model(inputs)
    losses = model.loss(out)
    metrics = model.get_metrics(out)
    torch.cuda.synchronize()
    ref_loss = float(g[case + "/nce_loss"])
    assert abs(float(losses["nce_loss"][1]) - ref_loss) / abs(ref_loss) < LOSS_TOL
    assert losses["nce_loss"][0] == 1.0
    assert out["vince_loss_dists"].shape == g[case + "/dists"].shape
    assert rel(out["vince_loss_dists"], g[case + "/dists"]) < LOSS_TOL
    assert rel(out["vince_loss_softmax_weights"], g[case + "/softmax_weights"]) < 5e-3
    assert abs(float(out["vince_loss_softmax_weight"]) - float(g[case + "/softmax_weight"])) < 5e-3 * max(
        float(g[case + "/softmax_weight"]), 1e-6)
    for name in ("nce_accuracy_mean", "cosine_sim", "cosine_sim_neg_max", "nce_softmax_weight_mean"):
        ref_v = float(g[case + "/metric_" + name])
        assert abs(float(metrics[name]) - ref_v) < 1e-3 * max(abs(ref_v), 1.0), name
    if self_cmp:
        ref_self = float(g[case + "/nce_loss_self"])
        assert abs(float(losses["nce_loss_self"][1]) - ref_self) / abs(ref_self) < LOSS_TOL
        assert rel(out["vince_loss_self_dists"], g[case + "/self_dists"]) < LOSS_TOL
    # fused backward: d(nce_loss [+ nce_loss_self])/d(embeddings) vs the reference's autograd result
    dq = model.embedding_gradients(out)
    torch.cuda.synchronize()
    err_dq = rel(dq, g[case + "/dq"])
    print("%s dq rel-L2 vs reference autograd: %.3e" % (case, err_dq))
    assert err_dq < 1e-3
    # the lazily materialised similarity matrix equals the reference's
    sims = out["vince_similarities"].materialize()
    assert rel(sims, g[case + "/similarities"]) < 1e-4
    # explicit-matrix API (loss_util.similarity_cross_entropy's literal signature)
    mask = torch.from_numpy(g[case + "/mask"]).to(DEV)
    ce = vince_b200.loss_util.similarity_cross_entropy(torch.from_numpy(g[case + "/similarities"]).to(DEV), T,
                                                       q.shape[0], 1, mask)
    torch.cuda.synchronize()
    assert abs(float(ce["dist"]) - ref_loss) / abs(ref_loss) < 1e-5
    assert rel(ce["dists"], g[case + "/dists"]) < 1e-5
    assert rel(ce["softmax_weights"], g[case + "/softmax_weights"]) < 1e-4


# ----------------------------------------------------------------------------------------------------------
# queue ring buffer: bit-exact vs the reference's StorageQueue
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["exact_multiple", "ragged", "bigger_than_queue"])
def test_queue_matches_reference_golden(golden, case):
    import vince_b200
    g = golden("queue.npz")
    init = torch.from_numpy(g[case + "/init"])
    K, D = init.shape
    q = vince_b200.StorageQueue(K, D, device=DEV)
    assert q.maxsize == K and len(q) == K and q.current_tail == 0 and q.full is False
    assert rel(q.vector_queue.norm(dim=1), torch.ones(K)) < 1e-6       # unit-norm random init
    q.load(init.to(DEV))
    for i, n in enumerate(g[case + "/sizes"]):
        items = torch.from_numpy(g["%s/items%d" % (case, i)]).to(DEV)
        q.enqueue(items, [i] * int(n), "src")
        torch.cuda.synchronize()
        assert torch.equal(q.dequeue()["queue_vectors"].cpu(), torch.from_numpy(g["%s/queue%d" % (case, i)])), i
        tail, full = g["%s/state%d" % (case, i)]
        assert q.current_tail == int(tail) and q.full == bool(full)
        # TF32 shadow stays coherent with the fp32 queue
        shadow = q.dequeue()["queue_vectors_tf32"]
        assert rel(shadow, q.vector_queue) < 3e-4
        assert bool(((shadow.view(torch.int32) & 0x1FFF) == 0).all())
    assert q.data_source_queue[q.current_tail - 1 if q.current_tail else K - 1] == "src"
    q.clear()
    assert q.current_tail == 0 and q.full is False


# ----------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: one full scoring step at 224x224 vs the reference's outputs
# ----------------------------------------------------------------------------------------------------------
def test_full_step_cfg0_matches_reference_golden(golden):
    import vince_b200
    g = golden("step_cfg0.npz")
    B, nf, K, D = [int(v) for v in g["cfg"]]
    T, m = [float(v) for v in g["T_m"]]
    args, model, sd = build_model("ResNet18", nf, B, K, D, T=T, seed=0)
    np.testing.assert_allclose(sd_checksum(sd), g["weights_checksum"], rtol=1e-9)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
    gen = torch.Generator().manual_seed(1234)
    data = torch.randn((B, 3, 224, 224), generator=gen)
    queue_data = torch.randn((B, 3, 224, 224), generator=gen)
    queue_init = F.normalize(torch.randn((K, D), generator=gen), dim=-1)
    np.testing.assert_allclose(checksum(data), g["data_checksum"], rtol=1e-9)
    np.testing.assert_allclose(checksum(queue_data), g["queue_data_checksum"], rtol=1e-9)
    np.testing.assert_allclose(checksum(queue_init), g["queue_init_checksum"], rtol=1e-9)
    queue = vince_b200.StorageQueue(K, D, device=DEV)
    queue.load(queue_init.to(DEV))
    queue.current_tail = K - 3
    batch = {"data": data.to(DEV), "queue_data": queue_data.to(DEV), "batch_types": ["images"], "batch_sizes": [B],
             "data_source": "synthetic", "num_frames": nf}
    # --- vince_solver.py:405-428 ---
    with injected_randperm([torch.from_numpy(g["perm_k"]), torch.from_numpy(g["perm_q"])]):
        queue_batches = qm(batch, shuffle=True)
        outputs = model.get_embeddings(batch, shuffle=True)
    output = outputs[0]
    output.update(queue.dequeue())
    output.update({"data_source": "synthetic", "num_frames": nf})
    output.update(queue_batches[0])
    output.update(model(output))
    loss = model.loss(output)
    metrics = model.get_metrics(output)
    # --- :497-499 ---
    queue.enqueue(output["queue_embeddings"], [None] * B, "synthetic")
    qm.vince_update(model)
    torch.cuda.synchronize()
    assert rel(output["embeddings"], g["embeddings"]) < EMB_TOL
    assert rel(output["queue_embeddings"], g["queue_embeddings"]) < EMB_TOL
    assert rel(output["extracted_features"], g["extracted_features"]) < EMB_TOL
    ref_loss = float(g["loss"])
    assert abs(float(loss["nce_loss"][1]) - ref_loss) / abs(ref_loss) < LOSS_TOL
    assert rel(output["vince_loss_dists"], g["dists"]) < LOSS_TOL
    for name in ("nce_accuracy_mean", "cosine_sim", "cosine_sim_neg_max", "nce_softmax_weight_mean"):
        ref_v = float(g["metric_" + name])
        assert abs(float(metrics[name]) - ref_v) < 2e-3 * max(abs(ref_v), 1.0), name
    tail, full = [int(v) for v in g["queue_state"]]
    assert queue.current_tail == tail and queue.full == bool(full)
    assert rel(queue.vector_queue[K - 3:], g["queue_tail_rows"]) < EMB_TOL
    assert rel(queue.vector_queue[:B], g["queue_head_rows"]) < EMB_TOL
    # rows that were not overwritten are bit-identical to the initial queue
    assert torch.equal(queue.vector_queue[B:K - 3].cpu(), queue_init[B:K - 3])
    np.testing.assert_allclose(sum(checksum(p) for p in qm.queue_network.vince_parameters()), g["ema_checksum"],
                               rtol=1e-6)


# ----------------------------------------------------------------------------------------------------------
# CUDA path vs the CPU oracle on seeded inputs (fp64 oracle = "truth")
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("backbone,B,H,passes,tol", [("ResNet18", 16, 96, 3, EMB_TOL), ("ResNet50", 8, 64, 3, EMB_TOL),
                                                    ("ResNet18", 16, 96, 1, 5e-2)])
def test_encoder_vs_fp64_oracle(backbone, B, H, passes, tol):
    args, model, sd = build_model(backbone, 4, B, 64, 128, seed=5, passes=passes)
    gen = torch.Generator().manual_seed(77)
    x = torch.randn((B, 3, H, H), generator=gen)
    ref = vo.get_embeddings(x.double(), vo.clone_state_dict(sd, torch.float64), backbone, True)
    r = model.get_embeddings({"data": x.to(DEV)})
    torch.cuda.synchronize()
    err = rel(r["embeddings"], ref["embeddings"])
    print("%s passes=%d embedding rel-L2 vs fp64 oracle: %.3e" % (backbone, passes, err))
    assert err < tol


def test_fused_update_equals_separate_calls():
    """VinceQueueModel.vince_update(model, enqueue=...) (one launch) == enqueue() + vince_update() (reference order)."""
    import vince_b200
    args, model, sd = build_model("ResNet18", 2, 8, 40, 32, seed=2)
    gen = torch.Generator().manual_seed(3)
    init = F.normalize(torch.randn((40, 32), generator=gen), dim=-1).to(DEV)
    keys = [F.normalize(torch.randn((8, 32), generator=gen), dim=-1).to(DEV) for _ in range(7)]
    res = []
    for fused in (False, True):
        qm = vince_b200.VinceQueueModel(args, model)
        qm.to(DEV)
        with torch.no_grad():
            for p in qm.queue_network.vince_parameters():
                p.mul_(0.5)
        queue = vince_b200.StorageQueue(40, 32, device=DEV)
        queue.load(init)
        queue.current_tail = 5
        for kk in keys:
            if fused:
                qm.vince_update(model, enqueue=(queue, kk, [None] * 8, "s"))
            else:
                queue.enqueue(kk, [None] * 8, "s")
                qm.vince_update(model)
        torch.cuda.synchronize()
        res.append((queue.vector_queue.clone(), queue.current_tail, queue.full,
                    [p.clone() for p in qm.queue_network.vince_parameters()]))
    assert torch.equal(res[0][0], res[1][0]) and res[0][1:3] == res[1][1:3]
    for a, b in zip(res[0][3], res[1][3]):
        assert torch.equal(a, b)


# ----------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]/[2] sizes)
# ----------------------------------------------------------------------------------------------------------
def test_infonce_full_size_vs_fp64_oracle():
    """B=256, K=65536, D=128, nf=4 (cfg1/cfg2 InfoNCE shape): fused kernel vs the fp64 oracle on the same bits."""
    from vince_b200 import ops
    gen = torch.Generator().manual_seed(99)
    B, K, D, nf = 256, 65536, 128, 4
    for T in (0.07, 0.2):
        q = F.normalize(torch.randn((B, D), generator=gen), dim=1)
        k = F.normalize(q + 0.8 * torch.randn((B, D), generator=gen), dim=1)
        queue = F.normalize(torch.randn((K, D), generator=gen), dim=1)
        losses, metrics, ex = vo.infonce(q.double(), k.double(), queue.double(), nf, T)
        qd, kd, qud = q.to(DEV), k.to(DEV), queue.to(DEV)
        qt = torch.empty_like(qud)
        ops.round_tf32(qud, qt)
        out = ops.infonce_fwd(qd, kd, qt, nf, T)
        torch.cuda.synchronize()
        sc = out["scalars"].cpu()
        assert abs(sc[0].item() - losses["nce_loss"].item()) / losses["nce_loss"].item() < 1e-4
        assert rel(out["dists"], ex["vince_loss_dists"].reshape(B, nf)) < 1e-4
        assert abs(sc[4].item() - metrics["cosine_sim_neg_max"].item()) < 1e-4
        assert abs(sc[3].item() - metrics["cosine_sim"].item()) < 1e-6
        # permutation invariance of the negatives: shuffling queue rows must not change the loss
        perm = torch.randperm(K, generator=gen)
        out2 = ops.infonce_fwd(qd, kd, qt[perm.to(DEV)].contiguous(), nf, T)
        torch.cuda.synchronize()
        assert abs(out2["scalars"][0].item() - sc[0].item()) < 1e-5 * abs(sc[0].item())
        assert torch.allclose(out2["neg_max"], out["neg_max"], atol=0, rtol=0)


def test_encoder_full_batch_eval_is_batch_separable():
    """Eval-mode BN makes frames independent: a B=256 224x224 ResNet-18 forward must equal, row for row, the same
    frames pushed through in two halves (exercises every tile-scheduling path at the benchmark's full size)."""
    args, model, sd = build_model("ResNet18", 4, 256, 64, 128, seed=1)
    model.eval()
    gen = torch.Generator().manual_seed(5)
    x = torch.randn((256, 3, 224, 224), generator=gen).to(DEV)
    full = model.get_embeddings({"data": x})["embeddings"]
    a = model.get_embeddings({"data": x[:128].contiguous()})["embeddings"]
    b = model.get_embeddings({"data": x[128:].contiguous()})["embeddings"]
    torch.cuda.synchronize()
    assert torch.isfinite(full).all()
    # not bit-exact: the tile-width / halo-vs-im2col choice depends on the batch, which changes the fp32 summation
    # order inside a convolution (channel-block-major vs tap-major); 1e-5 is ~2 orders below the parity tolerance
    err = rel(full, torch.cat((a, b)))
    print("full-batch vs two halves, eval mode: rel-L2 %.2e" % err)
    assert err < 1e-5
    assert rel(full.norm(dim=1), torch.ones(256)) < 1e-6


# ----------------------------------------------------------------------------------------------------------
# the BENCHMARKED shapes against the oracle: train-mode BN, B=256 frames of 224x224, injected shuffle.  These are the
# only cases that select the resident-weights tiles, the cta_group::2 N=64 layer-1 path at M=802816 and ResNet-50's
# 256-wide 1x1 tiles, i.e. exactly what bench.py times (BASELINE.json configs[1] / configs[2]).
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("backbone", ["ResNet18", "ResNet50"])
def test_encoder_full_size_train_mode_vs_oracle(backbone):
    B, H, D = 256, 224, 128
    args, model, sd = build_model(backbone, 4, B, 64, D, seed=11)
    gen = torch.Generator().manual_seed(2024)
    x = torch.randn((B, 3, H, H), generator=gen)
    perm = torch.randperm(B, generator=gen)
    ref_sd = vo.clone_state_dict(sd)
    with torch.no_grad():
        ref = vo.get_embeddings(x, ref_sd, backbone, True, shuffle_order=perm)          # fp32, the reference's arithmetic
    with injected_randperm([perm]):
        r = model.get_embeddings({"data": x.to(DEV)}, shuffle=True)
    torch.cuda.synchronize()
    errs = {k: rel(r[k], ref[k]) for k in ("embeddings", "prenorm_features", "extracted_features", "spatial_features")}
    print("%s B=%d %dx%d train-mode vs fp32 oracle: %s" % (backbone, B, H, H, {k: "%.2e" % v for k, v in errs.items()}))
    assert errs["embeddings"] < EMB_TOL and errs["prenorm_features"] < EMB_TOL
    assert errs["extracted_features"] < EMB_TOL and errs["spatial_features"] < EMB_TOL
    assert r["spatial_features"].shape == ref["spatial_features"].shape
    post = model.state_dict()
    for name in ("bn1.running_mean", "bn1.running_var", "layer1.0.bn1.running_var", "layer4.0.downsample.1.running_mean"):
        key = "feature_extractor.module.model." + name
        assert rel(post[key], ref_sd[key]) < 1e-4, name
    last_rv = [k for k in post if k.endswith("running_var")][-1]
    assert rel(post[last_rv], ref_sd[last_rv]) < 1e-3
    # uint8 HWC frames (bench.py's wire format) through the same full-size plan: bit-identical to host-normalised fp32
    from vince_b200 import ops
    x8 = torch.randint(0, 256, (B, H, H, 3), generator=gen, dtype=torch.uint8)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    xf = (((x8.permute(0, 3, 1, 2).float() / 255.0) - mean) / std).contiguous()
    snap = {k: v.clone() for k, v in model.state_dict().items()}
    outs = []
    for inp in (xf, x8):
        model.load_state_dict(snap)
        with injected_randperm([perm]):
            outs.append(model.get_embeddings({"data": inp.to(DEV)}, shuffle=True)["embeddings"])
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("backbone,B,H", [("ResNet50", 16, 96), ("ResNet18", 16, 64), ("ResNet50", 64, 224),
                                          ("ResNet50", 3, 72)])      # last: ragged tiles (972 / 243 / 75 / 27 rows)
def test_fused_epilogue_routes_equal_separate_bn_apply_bitwise(backbone, B, H):
    """The apply epilogue (BatchNorm scale/shift + residual + ReLU -> fp16 planes inside the convolution) performs the
    same fp32 operations in the same order as vince_bn_apply on the same accumulators, so (1) the train-mode
    statistics-pass + recompute-pass route for the wide 1x1 expansions and (2) the eval-mode folded-BatchNorm route must
    reproduce the raw-output + separate-bn_apply route bit for bit - embeddings AND running statistics."""
    args, model, sd = build_model(backbone, 4, B, 64, 128, seed=21)
    runner = model.feature_extractor.module.runner
    gen = torch.Generator().manual_seed(31)
    x = torch.randn((B, 3, H, H), generator=gen).to(DEV)
    snap = {k: v.clone() for k, v in model.state_dict().items()}
    res = {}
    saved = (runner.two_pass, runner.fold_eval, runner.tstats)
    for mode, train in (("separate", True), ("fused", True), ("fused_t", True), ("separate", False), ("fused", False)):
        model.load_state_dict(snap)
        model.train(train)
        runner.fold_eval = 0 if mode == "separate" else 1
        runner.two_pass = 0 if mode == "separate" else 2          # 2: every 1x1 expansion takes the two-pass route
        runner.tstats = 1 if mode == "fused_t" else 0
        out = model.get_embeddings({"data": x})
        torch.cuda.synchronize()
        res[(mode, train)] = (out["embeddings"].clone(), out["spatial_features"].clone(),
                              {k: v.clone() for k, v in model.state_dict().items() if "running" in k})
    runner.two_pass, runner.fold_eval, runner.tstats = saved
    for train in (True, False):
        a, b = res[("separate", train)], res[("fused", train)]
        assert torch.equal(a[0], b[0]), "embeddings differ (train=%s)" % train
        assert torch.equal(a[1], b[1]), "spatial features differ (train=%s)" % train
        for k in a[2]:
            assert torch.equal(a[2][k], b[2][k]), k
    assert not torch.equal(res[("fused", True)][0], res[("fused", False)][0])
    # transposed statistics pass: the same sums accumulated in another order (per channel over 256-pixel tiles, then
    # fp64) - equal to rounding of the fp32 coefficients, not bit for bit
    a, b = res[("separate", True)], res[("fused_t", True)]
    e_emb, e_sp = rel(b[0], a[0]), rel(b[1], a[1])
    e_run = max(rel(b[2][k], a[2][k]) for k in a[2] if not k.endswith("num_batches_tracked"))
    print("%s B=%d %d^2 transposed-statistics route vs separate bn_apply: embeddings %.2e spatial %.2e running stats %.2e"
          % (backbone, B, H, e_emb, e_sp, e_run))
    assert e_emb < 2e-6 and e_sp < 2e-6 and e_run < 1e-6


@pytest.mark.parametrize("M,N,K", [(802816, 256, 64), (6272, 512, 128), (98, 64, 64), (1000, 384, 192), (50176, 1024, 256)])
def test_transposed_statistics_pass_matches_fp64_sums(M, N, K):
    """vince_conv_fwd stats_only=2 (channels on the accumulator rows, in-register sums) against fp64 column sums of
    the same fp16x3 product and against the untransposed statistics pass; ragged M (TMA zero fill), N < 128, N not a
    multiple of 256 (single-CTA tiles) and the benchmark's largest shape."""
    from vince_b200 import ops
    gen = torch.Generator().manual_seed(5)
    x = torch.randn((M, K), generator=gen).to(DEV).relu_()
    w = (torch.randn((N, K), generator=gen) * 0.1).to(DEV)
    x_hi, x_lo = x.half(), (x - x.half().float()).half()
    w_hi, w_lo = w.half(), (w - w.half().float()).half()
    sums = {}
    for so in (1, 2):
        stats = torch.zeros((2 * N,), device=DEV, dtype=torch.float64)
        ops.conv_fwd(x_hi, x_lo, w_hi, w_lo, None, M, N, K, stats=stats, stats_only=so, alpha=0.5)
        torch.cuda.synchronize()
        sums[so] = stats
    xe, we = x_hi.double() + x_lo.double(), w_hi.double() + w_lo.double()
    ref_s, ref_q = torch.zeros(N, device=DEV, dtype=torch.float64), torch.zeros(N, device=DEV, dtype=torch.float64)
    for r0 in range(0, M, 65536):
        y = 0.5 * (xe[r0:r0 + 65536] @ we.t())
        ref_s += y.sum(0)
        ref_q += (y * y).sum(0)
    ref = torch.cat([ref_s, ref_q])
    # natural scale of each sum: |sum y| <= sqrt(M * sum y^2); the accumulators both passes read carry the tensor core's
    # own fp32 rounding (a few 1e-8 of the output rms per element, not zero-mean), so against fp64 the bar is 2e-4 of that
    # scale, while the two passes - which consume identical accumulators - must agree to fp32 summation accuracy
    scale = ref.abs() + 1e-3 * (ref_q * M).sqrt().repeat(2) + 1e-30
    for so in (1, 2):
        err = ((sums[so] - ref).abs() / scale).max().item()
        print("stats_only=%d M=%d N=%d K=%d: max rel err of the sums vs fp64 %.2e" % (so, M, N, K, err))
        assert err < 2e-4, (so, err)
    agree = ((sums[1] - sums[2]).abs() / scale).max().item()
    print("transposed vs untransposed statistics pass: %.2e" % agree)
    assert agree < 2e-6


def test_cfg4_resnet50_jigsaw_full_size_vs_oracle():
    """BASELINE.json configs[4] encoder side: ResNet-50 + jigsaw head, B=128 frames of 224x224 -> 1152 patches of 75x75
    (225-padded), per-row patch orders and the batch shuffle injected; EMBEDDINGS against the fp32 oracle."""
    B, H, D = 128, 224, 128
    args, model, sd = build_model("ResNet50", 4, B, 64, D, seed=13, jigsaw=True)
    gen = torch.Generator().manual_seed(44)
    x = torch.randn((B, 3, H, H), generator=gen)
    perm = torch.randperm(B, generator=gen)
    orders = torch.stack([torch.randperm(9, generator=gen) for _ in range(B)])
    with torch.no_grad():
        ref = vo.get_embeddings(x, vo.clone_state_dict(sd), "ResNet50", True, shuffle_order=perm, jigsaw=True,
                                jigsaw_orders=orders)
    with injected_randperm([perm]):
        r = model.get_embeddings({"data": x.to(DEV)}, jigsaw=True, shuffle=True, jigsaw_orders=orders)
    torch.cuda.synchronize()
    e1, e2 = rel(r["embeddings"], ref["embeddings"]), rel(r["prenorm_features"], ref["prenorm_features"])
    print("cfg4 R50+jigsaw B=%d: embeddings %.2e prenorm %.2e vs fp32 oracle" % (B, e1, e2))
    assert e1 < EMB_TOL and e2 < EMB_TOL
    assert r["embeddings"].shape == (B, D)


def test_jigsaw_uint8_and_nonsquare_pad_quirk():
    """(1) uint8 HWC frames through the jigsaw branch == host-normalised fp32 frames, bit for bit; (2) vince_model.py:
    145-146 pads BOTH axes by 3 - dim % 3 when either is not a multiple of 3 (a 48x50 frame becomes 51x51, 17x17
    patches): the stem-fused patchify must reproduce that geometry (oracle = F.pad + reshape)."""
    from vince_b200 import ops
    args, model, sd = build_model("ResNet18", 2, 4, 64, 128, seed=6, jigsaw=True)
    gen = torch.Generator().manual_seed(8)
    assert ops.jigsaw_patch_size(48, 50) == (17, 17) and ops.jigsaw_patch_size(48, 51) == (16, 17)
    assert ops.jigsaw_patch_size(224, 224) == (75, 75)
    for (H, W) in ((48, 50), (66, 66)):
        x8 = torch.randint(0, 256, (4, H, W, 3), generator=gen, dtype=torch.uint8)
        mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
        xf = (((x8.permute(0, 3, 1, 2).float() / 255.0) - mean) / std).contiguous()
        orders = torch.stack([torch.randperm(9, generator=gen) for _ in range(4)])
        snap = {k: v.clone() for k, v in model.state_dict().items()}
        outs = []
        for inp in (xf, x8):
            model.load_state_dict(snap)
            outs.append(model.get_embeddings({"data": inp.to(DEV)}, jigsaw=True, jigsaw_orders=orders))
        torch.cuda.synchronize()
        assert torch.equal(outs[0]["embeddings"], outs[1]["embeddings"])
        with torch.no_grad():
            ref = vo.get_embeddings(xf, vo.clone_state_dict(sd), "ResNet18", True, jigsaw=True, jigsaw_orders=orders)
        assert rel(outs[0]["embeddings"], ref["embeddings"]) < EMB_TOL, (H, W)


def test_queue_shadow_follows_out_of_band_writes():
    """code that writes StorageQueue.vector_queue directly (checkpoint restore, copy_) must not leave the
    loss running against a stale TF32 shadow."""
    import vince_b200
    gen = torch.Generator().manual_seed(1)
    q = vince_b200.StorageQueue(64, 32, device=DEV)
    new = F.normalize(torch.randn((64, 32), generator=gen), dim=-1).to(DEV)
    q.vector_queue.copy_(new)                               # torch-level write: bumps the version counter
    shadow = q.dequeue()["queue_vectors_tf32"]
    torch.cuda.synchronize()
    assert rel(shadow, new) < 3e-4
    q.vector_queue = (new * 0.5).contiguous()               # re-assignment: new storage
    shadow = q.dequeue()["queue_vectors_tf32"]
    assert rel(shadow, new * 0.5) < 3e-4
    q.enqueue(new[:8], [None] * 8, "s")                     # our own kernel keeps both coherent without a refresh
    assert not q._shadow_is_stale()
    assert rel(q.dequeue()["queue_vectors_tf32"][:8], new[:8]) < 3e-4


def test_loss_backward_fails_loudly_or_trains():
    """vince_solver.py:463-469 calls loss.backward(): the fused loss must carry a grad_fn and either run the encoder
    backward or raise a clear NotImplementedError - never autograd's opaque 'does not require grad'."""
    import vince_b200
    args, model, sd = build_model("ResNet18", 2, 4, 64, 128, seed=2)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
    gen = torch.Generator().manual_seed(3)
    batch = {"data": torch.randn((4, 3, 64, 64), generator=gen).to(DEV),
             "queue_data": torch.randn((4, 3, 64, 64), generator=gen).to(DEV), "batch_types": ["images"],
             "batch_sizes": [4], "data_source": "s", "num_frames": 2}
    queue = vince_b200.StorageQueue(64, 128, device=DEV)
    kb = qm(batch, shuffle=True)[0]
    out = model.get_embeddings(batch, shuffle=True)[0]
    out.update(queue.dequeue())
    out.update({"data_source": "s", "num_frames": 2})
    out.update(kb)
    out.update(model(out))
    loss = model.loss(out)["nce_loss"][1]
    assert loss.requires_grad and loss.grad_fn is not None
    if getattr(model, "supports_backward", False):
        loss.backward()
        assert all(p.grad is not None for p in model.embedding.parameters())
    else:
        with pytest.raises(NotImplementedError, match="query-encoder backward"):
            loss.backward()
    with torch.no_grad():
        assert not model.loss(out)["nce_loss"][1].requires_grad


# ----------------------------------------------------------------------------------------------------------
# training step: loss.backward() through the sm_100a backward kernels + SGD
# ----------------------------------------------------------------------------------------------------------
def _oracle_grads(sd, data, queue_data, queue_init, backbone, nf, T, perm_q, perm_k, dtype):
    q_sd = {}
    for k, v in sd.items():
        v = v.clone().to(dtype) if v.is_floating_point() else v.clone()
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
        q_sd[k] = v
    k_sd = vo.clone_state_dict(sd, dtype)
    with torch.no_grad():
        k = vo.get_embeddings(queue_data.to(dtype), k_sd, backbone, True, shuffle_order=perm_k)
    q = vo.get_embeddings(data.to(dtype), q_sd, backbone, True, shuffle_order=perm_q)
    losses, _, _ = vo.infonce(q["embeddings"], k["embeddings"], queue_init.to(dtype), nf, T)
    losses["nce_loss"].backward()
    return {k: v.grad.double() for k, v in q_sd.items() if v.is_floating_point() and v.grad is not None}


def _global_rel(a, b):
    num = sum(((a[k].double().cpu() - b[k]).norm() ** 2).item() for k in b)
    den = sum((b[k].norm() ** 2).item() for k in b)
    return (num / den) ** 0.5


@pytest.mark.parametrize("backbone,B,H", [("ResNet18", 32, 96), ("ResNet50", 16, 96)])
def test_backward_matches_oracle_autograd(backbone, B, H):
    """loss.backward() (vince_solver.py:463-469) against the oracle's autograd.  Truth = the fp64 oracle; the yardstick
    is the reference arithmetic itself: a random-init BatchNorm ResNet is so ill-conditioned at these sizes that the
    fp32 oracle's own gradients sit 3e-3 (ResNet-18) / 2e-2 (ResNet-50) away from the fp64 ones, so the kernels must be
    within 1e-3 OR within 2.5x of the fp32 reference's own distance from the truth (and every parameter the reference
    gives a gradient must get one, the unused torchvision `fc` none)."""
    import vince_b200
    nf, K, D, T = 2, 256, 128, 0.07
    args, model, sd = build_model(backbone, nf, B, K, D, T=T, seed=3)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
    g = torch.Generator().manual_seed(17)
    data, queue_data = torch.randn((B, 3, H, H), generator=g), torch.randn((B, 3, H, H), generator=g)
    queue_init = F.normalize(torch.randn((K, D), generator=g), dim=-1)
    perm_q, perm_k = torch.randperm(B, generator=g), torch.randperm(B, generator=g)
    queue = vince_b200.StorageQueue(K, D, device=DEV)
    queue.load(queue_init.to(DEV))
    batch = {"data": data.to(DEV), "queue_data": queue_data.to(DEV), "batch_types": ["images"], "batch_sizes": [B],
             "data_source": "s", "num_frames": nf}
    with injected_randperm([perm_k, perm_q]):
        kb = qm(batch, shuffle=True)[0]
        out = model.get_embeddings(batch, shuffle=True)[0]
    out.update(queue.dequeue())
    out.update({"data_source": "s", "num_frames": nf})
    out.update(kb)
    out.update(model(out))
    loss = model.loss(out)["nce_loss"][1]
    loss.backward()
    torch.cuda.synchronize()
    ref64 = _oracle_grads(sd, data, queue_data, queue_init, backbone, nf, T, perm_q, perm_k, torch.float64)
    ref32 = _oracle_grads(sd, data, queue_data, queue_init, backbone, nf, T, perm_q, perm_k, torch.float32)
    ours = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(ours) == set(ref64), set(ours) ^ set(ref64)
    assert "feature_extractor.module.model.fc.weight" not in ours
    e_ours, e_ref = _global_rel(ours, ref64), _global_rel(ref32, ref64)
    worst = max(((ours[k].double().cpu() - ref64[k]).norm() / ref64[k].norm()).item() for k in ref64)
    print("%s B=%d %dx%d gradients vs fp64 oracle: global rel-L2 %.2e (fp32 oracle autograd: %.2e), worst tensor %.2e"
          % (backbone, B, H, H, e_ours, e_ref, worst))
    assert e_ours < max(1e-3, 2.5 * e_ref)
    # a second backward without zero_grad accumulates (autograd semantics)
    g1 = {n: v.clone() for n, v in ours.items()}
    loss2 = model.loss(out)["nce_loss"][1]
    loss2.backward()
    torch.cuda.synchronize()
    name = "feature_extractor.module.model.layer1.0.conv1.weight"
    assert rel(dict(model.named_parameters())[name].grad, 2 * g1[name]) < 1e-5


def test_train_step_cfg0_matches_reference_golden(golden):
    """BASELINE.json configs[0] with the backward: gradients of the reference's OWN autograd (tests/golden/step_cfg0.npz,
    generated from the unmodified reference) for embedding.2.weight and conv1.weight."""
    import vince_b200
    g = golden("step_cfg0.npz")
    B, nf, K, D = [int(v) for v in g["cfg"]]
    T, m = [float(v) for v in g["T_m"]]
    args, model, sd = build_model("ResNet18", nf, B, K, D, T=T, seed=0)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
    gen = torch.Generator().manual_seed(1234)
    data = torch.randn((B, 3, 224, 224), generator=gen)
    queue_data = torch.randn((B, 3, 224, 224), generator=gen)
    queue_init = F.normalize(torch.randn((K, D), generator=gen), dim=-1)
    queue = vince_b200.StorageQueue(K, D, device=DEV)
    queue.load(queue_init.to(DEV))
    queue.current_tail = K - 3
    batch = {"data": data.to(DEV), "queue_data": queue_data.to(DEV), "batch_types": ["images"], "batch_sizes": [B],
             "data_source": "synthetic", "num_frames": nf}
    with injected_randperm([torch.from_numpy(g["perm_k"]), torch.from_numpy(g["perm_q"])]):
        queue_batches = qm(batch, shuffle=True)
        outputs = model.get_embeddings(batch, shuffle=True)
    output = outputs[0]
    output.update(queue.dequeue())
    output.update({"data_source": "synthetic", "num_frames": nf})
    output.update(queue_batches[0])
    output.update(model(output))
    loss = model.loss(output)["nce_loss"]
    (loss[0] * loss[1]).backward()
    torch.cuda.synchronize()
    ge = model.embedding[2].weight.grad
    gc = model.feature_extractor.module.model.conv1.weight.grad
    e_row = rel(ge[0], g["grad_embedding2_weight_row0"])
    c_e, c_c = checksum(ge), checksum(gc)
    print("cfg0 gradients vs the reference's autograd: embedding.2.weight row0 rel %.2e, checksums %s / %s (reference %s / %s)"
          % (e_row, c_e, c_c, g["grad_embedding2_weight_checksum"], g["grad_conv1_checksum"]))
    # one 512-element row of a B=8 step is dominated by the fp32 noise of the reference's own autograd (3e-3 .. 2e-2 away
    # from an fp64 evaluation at these sizes, test_backward_matches_oracle_autograd); the sums of |g| are the tight check
    assert e_row < 2e-2
    np.testing.assert_allclose(c_e[2], g["grad_embedding2_weight_checksum"][2], rtol=5e-3)       # sum |g|
    np.testing.assert_allclose(c_c[2], g["grad_conv1_checksum"][2], rtol=2e-2)


def _bare_solver(VinceSolver, args, model, queue_model, queue, batches):
    """A VinceSolver with exactly the state run_train_iteration reads (vince_solver.py:386-518), built without its
    __init__ (datasets, loaders, tensorboard, CIFAR): models, queue, the reference's SGD recipe (:252-256), meters."""
    import collections
    from dg_util.python_utils.average_meter import RollingAverageMeter
    s = object.__new__(VinceSolver)
    s.args = args
    s.model, s.queue_model, s.vince_queue = model, queue_model, queue
    s.use_apex = False
    # (the recipe of :252-256 with a small base_lr: at B = 8 the gradient itself carries 1e-3 .. 2e-2 of fp32 noise in the
    #  reference's own autograd, which a large step would turn into a visible difference of the NEXT iteration's loss)
    s.optimizer = torch.optim.SGD([{"params": model.parameters(), "initial_lr": 0.003}], lr=0.003, weight_decay=0.0001,
                                  momentum=0.9)
    s.time_meters = collections.defaultdict(RollingAverageMeter)
    s.loss_meters = collections.defaultdict(RollingAverageMeter)
    s.metric_meters = collections.defaultdict(RollingAverageMeter)
    s.train_logger = None
    s.drawn_this_epoch = True
    s.logger_iteration, s.iteration = 1, 0            # (full_name / model_name are properties derived from the classes)
    it = iter(batches)
    s.get_batch = lambda: (next(it), None)
    return s


def test_reference_solver_runs_unchanged_on_vince_b200_classes():
    """The drop-in claim, executed: the reference's OWN VinceSolver.run_train_iteration (solvers/vince_solver.py:386-518,
    unmodified source staged under oracle/_ref) drives vince_b200's VinceModel / VinceQueueModel / StorageQueue on the GPU
    for two training iterations - forward, loss_dict / metrics handling, optimizer.zero_grad(), loss.backward(),
    torch.optim.SGD.step(), enqueue, vince_update - and is compared with the same two iterations of the same solver
    code over the reference's own classes on the CPU (same weights, inputs, shuffles, queue)."""
    import types
    import vince_b200
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_loader
    if not (ref_loader.reference_available() and ref_loader.solver_available()):
        pytest.skip("reference solver not staged (oracle/build_ref.py)")
    VinceSolver = ref_loader.load_reference_solver()
    ref = ref_loader.load_reference()
    B, nf, K, D, H = 8, 2, 64, 128, 64
    gen = torch.Generator().manual_seed(77)
    batches_cpu = [{"data": torch.randn((B, 3, H, H), generator=gen), "queue_data": torch.randn((B, 3, H, H), generator=gen),
                    "batch_types": ["images"], "batch_sizes": [B], "num_frames": [nf], "data_source": ["synthetic"],
                    "queue_data_cpu": [None] * B} for _ in range(2)]      # per-type lists, as get_batch builds them (:365)
    perms = [torch.randperm(B, generator=gen) for _ in range(4)]
    queue_init = F.normalize(torch.randn((K, D), generator=gen), dim=-1)
    extra = dict(save_frequency=10 ** 9, log_frequency=10 ** 9, base_lr=0.003)

    # ---- the reference's classes on the CPU ----
    rargs = ref_loader.make_args(backbone="ResNet18", num_frames=nf, batch_size=B, queue_size=K, embedding_size=D)
    for k_, v_ in extra.items():
        setattr(rargs, k_, v_)
    sd = vo.make_state_dict("ResNet18", D, seed=3)
    rmodel = ref.VinceModel(rargs)
    rmodel.load_state_dict(sd, strict=True)
    rmodel.train()
    rqm = ref.VinceQueueModel(rargs, rmodel)
    rqm.train()
    rqueue = ref.StorageQueue(K, D, device="cpu")
    rqueue.vector_queue.copy_(queue_init)
    rs = _bare_solver(VinceSolver, rargs, rmodel, rqm, rqueue, [dict(b) for b in batches_cpu])
    ref_losses = []
    with injected_randperm([p.clone() for p in perms]):
        for _ in range(2):
            rs.run_train_iteration()
            ref_losses.append(rs.loss_meters["nce_loss"].history[-1])

    # ---- vince_b200's classes on the GPU, same solver code ----
    args, model, _ = build_model("ResNet18", nf, B, K, D, seed=3)
    for k_, v_ in extra.items():
        setattr(args, k_, v_)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
    queue = vince_b200.StorageQueue(K, D, device=DEV)
    queue.load(queue_init.to(DEV))
    batches = [{k_: (v_.to(DEV) if isinstance(v_, torch.Tensor) else v_) for k_, v_ in b.items()} for b in batches_cpu]
    s = _bare_solver(VinceSolver, args, model, qm, queue, batches)
    losses = []
    with injected_randperm([p.clone() for p in perms]):
        for _ in range(2):
            s.run_train_iteration()
            losses.append(s.loss_meters["nce_loss"].history[-1])
    torch.cuda.synchronize()
    print("reference solver loop: losses ours %s / reference %s" % (["%.5f" % v for v in losses], ["%.5f" % v for v in ref_losses]))
    assert s.iteration == rs.iteration == 2 * B and queue.current_tail == rqueue.current_tail == 2 * B
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) / abs(b) < 2e-3
    for name in ("nce_accuracy_mean", "cosine_sim", "cosine_sim_neg_max"):
        assert abs(s.metric_meters[name].val - rs.metric_meters[name].val) < 5e-3, name
    assert rel(queue.vector_queue.cpu(), rqueue.vector_queue) < 1e-3
    post, rpost = model.state_dict(), rmodel.state_dict()
    keys = [k_ for k_ in rpost if rpost[k_].is_floating_point()]
    # all parameters / buffers together (weights dominate), and the two-step UPDATE of every tensor on its own: the update
    # is lr * (momentum-filtered gradient), whose fp32 noise at B = 8 is 1e-3 .. 2e-2 in the reference's own autograd
    # (test_backward_matches_oracle_autograd), so that bar is a sanity bound, not a precision claim
    num = sum(float(((post[k_].cpu().double() - rpost[k_].double()) ** 2).sum()) for k_ in keys)
    den = sum(float((rpost[k_].double() ** 2).sum()) for k_ in keys)
    glob = (num / den) ** 0.5
    upd = []
    for k_ in keys:
        if k_ in sd and sd[k_].is_floating_point() and "running" not in k_:
            du, dr = post[k_].cpu().double() - sd[k_].double(), rpost[k_].double() - sd[k_].double()
            if float(dr.norm()) > 0:
                upd.append((float((du - dr).norm() / dr.norm()), k_))
    worst = max(upd)
    print("after two SGD steps: all parameters + buffers rel-L2 %.2e vs the reference; worst per-tensor update error %.2e (%s)"
          % (glob, worst[0], worst[1]))
    assert glob < 5e-4 and worst[0] < 2e-1          # (measured 1.2e-4 and 7.3e-2: BatchNorm gammas of a B = 8 step)
    for name in ("bn1.running_mean", "layer4.1.bn2.running_var"):
        k_ = "feature_extractor.module.model." + name
        assert rel(post[k_].cpu(), rpost[k_]) < 1e-3, name
    qpost, rqpost = qm.state_dict(), rqm.state_dict()
    qnum = sum(float(((qpost[k_].cpu().double() - rqpost[k_].double()) ** 2).sum()) for k_ in rqpost if rqpost[k_].is_floating_point())
    qden = sum(float((rqpost[k_].double() ** 2).sum()) for k_ in rqpost if rqpost[k_].is_floating_point())
    assert (qnum / qden) ** 0.5 < 5e-4         # key encoder after two EMA updates


def test_fused_sgd_matches_torch_sgd_and_trains():
    """vince_b200.optim.FusedSGD == torch.optim.SGD(momentum=0.9, weight_decay=1e-4) (vince_solver.py:252-256) over three
    steps on identical gradients; then a few real training steps must lower the loss on a fixed batch."""
    import vince_b200
    from vince_b200.optim import FusedSGD
    gen = torch.Generator().manual_seed(5)
    shapes = [(64, 3, 7, 7), (64,), (128, 64, 3, 3), (1000, 512), (17,)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=gen).to(DEV)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedSGD(pa, lr=0.03, momentum=0.9, weight_decay=1e-4)
    ob = torch.optim.SGD(pb, lr=0.03, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=gen).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
        for group in oa.param_groups:
            group["lr"] *= 0.5                           # solver_runner.py:36-43 rewrites param_group["lr"]
        for group in ob.param_groups:
            group["lr"] *= 0.5
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert rel(a, b) < 1e-6
        assert rel(oa.state[a]["momentum_buffer"], ob.state[b]["momentum_buffer"]) < 1e-6
    # end to end: the solver's loop (zero_grad, backward, step, enqueue, EMA) lowers the loss on a fixed batch
    B, nf, K, D = 16, 2, 256, 128
    args, model, sd = build_model("ResNet18", nf, B, K, D, seed=9)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
    opt = FusedSGD(model.parameters(), lr=0.003, momentum=0.9, weight_decay=1e-4)
    queue = vince_b200.StorageQueue(K, D, device=DEV)
    x = torch.randn((B, 3, 64, 64), generator=gen).to(DEV)
    batch = {"data": x, "queue_data": (x + 0.05 * torch.randn(x.shape, generator=gen).to(DEV)).contiguous(),
             "batch_types": ["images"], "batch_sizes": [B], "data_source": "s", "num_frames": nf}
    losses = []
    for it in range(6):
        kb = qm(batch, shuffle=True)[0]
        out = model.get_embeddings(batch, shuffle=True)[0]
        out.update(queue.dequeue())
        out.update({"data_source": "s", "num_frames": nf})
        out.update(kb)
        out.update(model(out))
        ld = model.loss(out)
        loss = ld["nce_loss"][0] * ld["nce_loss"][1]
        opt.zero_grad()
        loss.backward()
        opt.step()
        qm.vince_update(model)                       # the queue is left alone: a fixed batch against fixed negatives
        losses.append(float(loss))
    print("training losses on a fixed batch:", ["%.4f" % v for v in losses])
    assert min(losses[1:]) < losses[0] - 0.05 and losses[-1] < losses[0] and all(np.isfinite(losses))
    # fp16 range check on these SGD-stepped (no longer random-init) weights: every activation plane of a train- and an
    # eval-mode forward stays clear of the +-65504 saturation bound; with one BatchNorm gamma blown up the counter and the
    # warning fire (VINCE_B200_CHECK_SATURATION=1 does the same in any run)
    import warnings
    runner = model.feature_extractor.module.runner
    runner.check_saturation = 1
    for train in (True, False):
        model.train(train)
        with torch.no_grad():
            model.get_embeddings({"data": x})
        assert runner.saturated == 0, runner.saturated
    with torch.no_grad():
        model.feature_extractor.module.model.layer1[0].bn1.weight.fill_(3e5)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        with torch.no_grad():
            model.get_embeddings({"data": x})
    print("saturation check: %d clamped activation values after blowing up one BatchNorm gamma" % runner.saturated)
    assert runner.saturated > 0 and any("saturation bound" in str(w.message) for w in caught)
    runner.check_saturation = 0


def test_knn_matches_sklearn_kdtree():
    """vince_solver.py:651-693: KDTree(features).query(k=11), drop self, mode of the neighbour labels, accuracy."""
    import scipy.stats
    from sklearn.neighbors import KDTree
    from vince_b200 import knn
    gen = torch.Generator().manual_seed(12)
    n, D, C = 3001, 128, 10
    centers = F.normalize(torch.randn((C, D), generator=gen), dim=1)
    labels = torch.randint(0, C, (n,), generator=gen)
    feats = F.normalize(centers[labels] + 0.35 * torch.randn((n, D), generator=gen), dim=1)
    nbr, dist, pred = knn.knn_classify(feats.to(DEV), labels.to(DEV), k=10)
    torch.cuda.synchronize()
    kdt = KDTree(feats.numpy(), leaf_size=40, metric="euclidean")
    ref_d, ref_i = kdt.query(feats.numpy(), k=11)
    ref_i, ref_d = ref_i[:, 1:], ref_d[:, 1:]
    ref_pred = scipy.stats.mode(labels.numpy()[ref_i], axis=1)[0].reshape(-1)
    ref_acc = float(np.mean(ref_pred == labels.numpy()))
    same_sets = np.mean([set(a) == set(b) for a, b in zip(nbr.cpu().numpy(), ref_i)])
    assert same_sets > 0.999, same_sets                         # fp32 vs fp64 distances may swap exact near-ties
    np.testing.assert_allclose(dist.cpu().numpy(), ref_d, rtol=1e-4, atol=1e-6)
    agree = float((pred.cpu().numpy() == ref_pred).mean())
    acc = float(knn.knn_accuracy(feats.to(DEV), labels.to(DEV), k=10))
    print("kNN: neighbour sets equal %.4f, predictions agree %.4f, accuracy %.4f (sklearn %.4f)" % (same_sets, agree, acc, ref_acc))
    assert agree > 0.999 and abs(acc - ref_acc) < 2e-3
    # end to end through the eval-mode encoder (folded BatchNorm), uint8 frames and the reference's float dataset tensor
    args, model, sd = build_model("ResNet18", 4, 64, 64, 128, seed=3)
    model.eval()
    imgs8 = torch.randint(0, 256, (96, 32, 32, 3), generator=gen, dtype=torch.uint8)
    labs = torch.randint(0, 10, (96,), generator=gen)
    acc8, feats8 = knn.knn_eval(model, imgs8, labs.to(DEV), batch_size=40)
    accf, featsf = knn.knn_eval(model, imgs8.permute(0, 3, 1, 2).float() / 255.0, labs.to(DEV), batch_size=40)
    torch.cuda.synchronize()
    assert feats8.shape == (96, 128) and 0.0 <= float(acc8) <= 1.0
    assert rel(feats8, featsf) < 1e-5


def test_multi_gpu_allgather_and_shard_parity():
    """Collected multi-GPU parity: skips below 2 visible GPUs; otherwise launches
    tests/multigpu_check.py under torchrun on every visible GPU (bit-exact all-gather + enqueue vs the oracle's
    enqueue(cat(keys_0..keys_{R-1})), per-shard step parity with a shared queue snapshot)."""
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    print(out.stdout[-1500:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert "multigpu_check world=%d: PASS" % n in out.stdout


def test_encoder_overlap_is_bitwise_equal_to_serial():
    """The two-stream schedule (key encoder on the caller's stream, query encoder on a side stream, joined before
    get_embeddings returns) must not change a single bit, and must only trigger for the batch the fork point saw."""
    import vince_b200
    from vince_b200 import vince_model as vm
    args, model, sd = build_model("ResNet18", 2, 8, 64, 128, seed=3)
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
    gen = torch.Generator().manual_seed(11)
    data = torch.randn((8, 3, 96, 96), generator=gen).to(DEV)
    queue_data = torch.randn((8, 3, 96, 96), generator=gen).to(DEV)
    batch = {"data": data, "queue_data": queue_data, "batch_types": ["images"], "batch_sizes": [8]}
    snap = {k: v.clone() for k, v in model.state_dict().items()}
    snap_k = {k: v.clone() for k, v in qm.state_dict().items()}
    results = []
    for overlap in ("0", "1", "1"):
        model.load_state_dict(snap)
        qm.load_state_dict(snap_k)
        os.environ["VINCE_B200_OVERLAP"] = overlap
        try:
            perms = [torch.randperm(8, generator=torch.Generator().manual_seed(5)) for _ in range(2)]
            with injected_randperm(perms):
                kb = qm(batch, shuffle=True)[0]
                qb = model.get_embeddings(batch, shuffle=True)[0]
            # consumed on the caller's stream with NO device-wide synchronisation in between: the join must be enough
            out = (qb["embeddings"] + 0).cpu(), (kb["queue_embeddings"] + 0).cpu(), (qb["spatial_features"] + 0).cpu()
        finally:
            os.environ.pop("VINCE_B200_OVERLAP", None)
        results.append(out)
    for a, b in zip(results[0], results[1]):
        assert torch.equal(a, b)
    for a, b in zip(results[0], results[2]):
        assert torch.equal(a, b)
    # the fork point is honoured only by the call that immediately follows, for the very tensor it saw, unmodified
    qm(batch, shuffle=False)
    vm._CALLS[0] += 1
    assert vm._take_fork(data) is not None
    qm(batch, shuffle=False)
    vm._CALLS[0] += 1
    other = {"data": data.clone(), "batch_types": ["images"], "batch_sizes": [8]}
    assert vm._take_fork(other["data"]) is None
    qm(batch, shuffle=False)
    data.add_(0.0)                      # bumps the version counter
    vm._CALLS[0] += 1
    assert vm._take_fork(data) is None
    qm(batch, shuffle=False)
    vm._CALLS[0] += 2                   # some other encoder-level call came in between
    assert vm._take_fork(data) is None
    torch.cuda.synchronize()


def test_uint8_hwc_input_equals_normalised_fp32_input():
    """raw uint8 HWC frames with ToTensor(scale=255) + Normalize (utils/transforms.py:89-101) fused
    into the stem packing must give bit-identical results to feeding the frames normalised on the host."""
    from vince_b200 import ops
    args, model, sd = build_model("ResNet18", 2, 6, 64, 128, seed=4)
    gen = torch.Generator().manual_seed(21)
    for (H, W) in ((96, 96), (75, 51)):
        x8 = torch.randint(0, 256, (6, H, W, 3), generator=gen, dtype=torch.uint8)
        mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
        xf = ((x8.permute(0, 3, 1, 2).float() / 255.0) - mean) / std          # ToTensor(scale=255) + Normalize
        perm = torch.randperm(6, generator=gen)
        snap = {k: v.clone() for k, v in model.state_dict().items()}
        outs = []
        for x in (xf.contiguous(), x8):
            model.load_state_dict(snap)
            with injected_randperm([perm]):
                outs.append(model.get_embeddings({"data": x.to(DEV)}, shuffle=True))
        torch.cuda.synchronize()
        for key in ("embeddings", "extracted_features", "spatial_features"):
            assert torch.equal(outs[0][key], outs[1][key]), key


def test_batch_prefetcher_orders_copies_and_compute():
    from vince_b200.prefetch import BatchPrefetcher
    pf = BatchPrefetcher(DEV, depth=2)
    host = [{"data": torch.full((1 << 20,), float(i)).pin_memory(), "queue_data": torch.full((4,), -float(i)).pin_memory(),
             "num_frames": 4} for i in range(5)]
    got = []
    pf.submit(host[0])
    for i in range(5):
        if i + 1 < 5:
            pf.submit(host[i + 1])
        b = pf.next()
        assert b["num_frames"] == 4 and b["queue_data_cpu"] is host[i]["queue_data"]
        got.append((b["data"].sum() / b["data"].numel(), b["queue_data"].sum()))     # async compute on the batch
        pf.release(b)
    with pytest.raises(RuntimeError):
        pf.next()
    torch.cuda.synchronize()
    for i, (m, q) in enumerate(got):
        assert float(m) == float(i) and float(q) == -4.0 * i
    assert pf.h2d_bytes == (1 << 20) * 4 + 16


def test_fp16_split_saturates_instead_of_overflowing():
    """Operand planes are fp16 (hi, lo): values beyond +-65504 must saturate, never become inf/NaN."""
    from vince_b200 import ops
    x = torch.tensor([0.0, 1.0, -3.14159, 70000.0, -1e9, 6e-8, 1e-3, 65504.0], device=DEV)
    hi = torch.empty_like(x, dtype=torch.float16)
    lo = torch.empty_like(hi)
    ops.split_f16(x, hi, lo)
    torch.cuda.synchronize()
    assert torch.isfinite(hi.float()).all() and torch.isfinite(lo.float()).all()
    back = hi.double() + lo.double()
    small = x.abs() <= 65504
    assert ((back[small] - x[small].double()).abs() <= x[small].double().abs() * 2.0 ** -22 + 6e-8).all()
    assert float(back[3]) == 70000.0 and float(hi[4]) == -65504.0


def test_ops_reject_cpu_tensors():
    from vince_b200 import ops
    with pytest.raises(RuntimeError):
        ops.l2_normalize(torch.zeros(4, 8), torch.zeros(4, 8))
