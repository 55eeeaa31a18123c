"""CPU: the oracle (oracle/vince_oracle.py) against the golden vectors generated from the UNMODIFIED reference
(tests/golden/*.npz, made by oracle/make_golden.py).  This is what pins the oracle; the GPU tests then compare the
CUDA path with both."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vince_oracle as vo


def rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def checksum(t):
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder(97.0) + 1.0
    return np.array([t.sum().item(), (t * w).sum().item(), t.abs().sum().item()])


def sd_checksum(sd):
    acc = np.zeros(3)
    for v in sd.values():
        if v.is_floating_point():
            acc += checksum(v)
    return acc


@pytest.mark.parametrize("case", ["ibc_nf4", "ibc_nf2", "ibc_nf1", "ibc_self_nf4", "moco", "ibc_short_batch"])
def test_infonce_oracle_matches_reference(golden, case):
    g = golden("infonce.npz")
    B, K, D, nf, ibc, self_cmp, Bact = [int(v) for v in g[case + "/cfg"]]
    T, T_self = [float(v) for v in g[case + "/T"]]
    q, k, queue = (torch.from_numpy(g[case + "/" + n]) for n in ("q", "k", "queue"))
    losses, metrics, ex = vo.infonce(q, k, queue, nf, T, bool(ibc), bool(self_cmp), T_self)
    assert torch.equal(ex["vince_similarities_mask"], torch.from_numpy(g[case + "/mask"]))
    assert rel(ex["vince_similarities"], g[case + "/similarities"]) < 1e-6
    assert rel(ex["vince_loss_dists"], g[case + "/dists"]) < 1e-5
    assert abs(float(losses["nce_loss"]) - float(g[case + "/nce_loss"])) < 1e-5 * abs(float(g[case + "/nce_loss"]))
    assert rel(ex["vince_loss_softmax_weights"], g[case + "/softmax_weights"]) < 1e-4
    for name in ("nce_accuracy_mean", "cosine_sim", "cosine_sim_neg_max", "nce_softmax_weight_mean"):
        assert abs(float(metrics[name]) - float(g[case + "/metric_" + name])) < 1e-5, name
    if self_cmp:
        assert abs(float(losses["nce_loss_self"]) - float(g[case + "/nce_loss_self"])) < 1e-5
    # gradient of the loss wrt the queries (what a fused backward has to reproduce)
    q2 = q.clone().requires_grad_(True)
    l2, _, _ = vo.infonce(q2, k, queue, nf, T, bool(ibc), bool(self_cmp), T_self)
    sum(l2.values()).backward()
    assert rel(q2.grad, g[case + "/dq"]) < 1e-4


@pytest.mark.parametrize("case", ["exact_multiple", "ragged", "bigger_than_queue"])
def test_queue_oracle_matches_reference(golden, case):
    g = golden("queue.npz")
    q = vo.StorageQueue(*g[case + "/init"].shape, init=torch.from_numpy(g[case + "/init"]))
    for i, n in enumerate(g[case + "/sizes"]):
        q.enqueue(torch.from_numpy(g["%s/items%d" % (case, i)]))
        assert torch.equal(q.vector_queue, torch.from_numpy(g["%s/queue%d" % (case, i)]))
        tail, full = g["%s/state%d" % (case, i)]
        assert q.current_tail == int(tail) and q.full == bool(full)


@pytest.mark.parametrize("case,backbone", [("r18_64", "ResNet18"), ("r50_64", "ResNet50"), ("r18_odd", "ResNet18")])
def test_encoder_oracle_matches_reference(golden, case, backbone):
    g = golden("encoder.npz")
    B, nf, H, W, D, seed, shuffle = [int(v) for v in g[case + "/cfg"]]
    sd = vo.make_state_dict(backbone, D, seed=seed)
    np.testing.assert_allclose(sd_checksum(sd), g[case + "/weights_checksum"], rtol=1e-9)
    x = torch.from_numpy(g[case + "/x"])
    perm = torch.from_numpy(g[case + "/perm"]) if shuffle else None
    with torch.no_grad():
        r = vo.get_embeddings(x, sd, backbone, True, shuffle_order=perm)
    tol = 2e-4 if backbone == "ResNet50" else 2e-5          # fp32 summation-order noise, amplified in random-init R50
    assert rel(r["embeddings"], g[case + "/embeddings"]) < tol
    assert rel(r["extracted_features"], g[case + "/extracted_features"]) < 5 * tol
    assert rel(r["spatial_features"], g[case + "/spatial_features"]) < 5 * tol
    assert rel(sd["feature_extractor.module.model.bn1.running_var"], g[case + "/bn1_running_var"]) < 1e-5
    assert int(sd["feature_extractor.module.model.bn1.num_batches_tracked"]) == int(g[case + "/bn1_num_batches"])
    with torch.no_grad():
        r_eval = vo.get_embeddings(x, sd, backbone, False)
    assert rel(r_eval["embeddings"], g[case + "/eval_embeddings"]) < tol
    # EMA of a perturbed query encoder onto a key copy (reference: VinceQueueModel.param_update)
    sd0 = vo.make_state_dict(backbone, D, seed=seed)          # key copy was taken after the forward: BN stats differ,
    key = vo.clone_state_dict(sd0)                            # but only parameters are averaged
    names = vo.vince_parameter_names(sd0)
    for i, n in enumerate(names):
        sd0[n].add_(0.01 * ((i % 7) - 3))
    vo.param_update(key, sd0, 0.999, names)
    n_tensors, n_elems = [int(v) for v in g[case + "/ema_n_tensors"]]
    assert len(names) == n_tensors and sum(key[n].numel() for n in names) == n_elems
    np.testing.assert_allclose(sum(checksum(key[n]) for n in names), g[case + "/ema_checksum"], rtol=1e-6)
    assert rel(key["embedding.2.bias"], g[case + "/ema_embedding2_bias"]) < 1e-6


def test_jigsaw_oracle_matches_reference(golden):
    g = golden("jigsaw.npz")
    case = "r18_jigsaw"
    B, nf, H, W, D, seed = [int(v) for v in g[case + "/cfg"]]
    sd = vo.make_state_dict("ResNet18", D, jigsaw=True, seed=seed)
    np.testing.assert_allclose(sd_checksum(sd), g[case + "/weights_checksum"], rtol=1e-9)
    with torch.no_grad():
        r = vo.get_embeddings(torch.from_numpy(g[case + "/x"]), sd, "ResNet18", True,
                              shuffle_order=torch.from_numpy(g[case + "/perm"]), jigsaw=True,
                              jigsaw_orders=torch.from_numpy(g[case + "/orders"]))
    assert rel(r["embeddings"], g[case + "/embeddings"]) < 2e-5
    assert rel(r["prenorm_features"], g[case + "/prenorm_features"]) < 2e-5


def test_full_step_cfg0_oracle_matches_reference(golden):
    """BASELINE.json configs[0] (ResNet18, 2 views, batch 8, K=1024, D=128, 224x224) - one scoring step."""
    g = golden("step_cfg0.npz")
    B, nf, K, D = [int(v) for v in g["cfg"]]
    T, m = [float(v) for v in g["T_m"]]
    sd = vo.make_state_dict("ResNet18", D, seed=0)
    np.testing.assert_allclose(sd_checksum(sd), g["weights_checksum"], rtol=1e-9)
    gen = torch.Generator().manual_seed(1234)
    data = torch.randn((B, 3, 224, 224), generator=gen)
    queue_data = torch.randn((B, 3, 224, 224), generator=gen)
    queue_init = F.normalize(torch.randn((K, D), generator=gen), dim=-1)
    np.testing.assert_allclose(checksum(data), g["data_checksum"], rtol=1e-9)
    np.testing.assert_allclose(checksum(queue_init), g["queue_init_checksum"], rtol=1e-9)
    q_sd, k_sd = vo.clone_state_dict(sd), vo.clone_state_dict(sd)
    queue = vo.StorageQueue(K, D, init=queue_init)
    queue.current_tail = K - 3
    out = vo.train_step(data, queue_data, q_sd, k_sd, queue, "ResNet18", nf, T, m,
                        shuffle_q=torch.from_numpy(g["perm_q"]), shuffle_k=torch.from_numpy(g["perm_k"]))
    assert rel(out["query"]["embeddings"], g["embeddings"]) < 2e-5
    assert rel(out["key"]["embeddings"], g["queue_embeddings"]) < 2e-5
    assert abs(float(out["losses"]["nce_loss"]) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    for name in ("nce_accuracy_mean", "cosine_sim", "cosine_sim_neg_max", "nce_softmax_weight_mean"):
        assert abs(float(out["metrics"][name]) - float(g["metric_" + name])) < 1e-4, name
    tail, full = [int(v) for v in g["queue_state"]]
    assert queue.current_tail == tail and queue.full == bool(full)
    assert rel(queue.vector_queue[K - 3:], g["queue_tail_rows"]) < 2e-5
    np.testing.assert_allclose(checksum(queue.vector_queue), g["queue_checksum"], rtol=1e-5)
    names = vo.vince_parameter_names(sd)
    np.testing.assert_allclose(sum(checksum(k_sd[n]) for n in names), g["ema_checksum"], rtol=1e-6)
    # gradients of the reference's loss (checks the oracle's graph end to end)
    q_sd2 = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
             for k, v in sd.items()}
    with torch.no_grad():
        kk = vo.get_embeddings(queue_data, vo.clone_state_dict(sd), "ResNet18", True, shuffle_order=torch.from_numpy(g["perm_k"]))
    qq = vo.get_embeddings(data, q_sd2, "ResNet18", True, shuffle_order=torch.from_numpy(g["perm_q"]))
    q0 = vo.StorageQueue(K, D, init=queue_init)
    losses, _, _ = vo.infonce(qq["embeddings"], kk["embeddings"], q0.vector_queue, nf, T)
    losses["nce_loss"].backward()
    np.testing.assert_allclose(checksum(q_sd2["embedding.2.weight"].grad), g["grad_embedding2_weight_checksum"], rtol=2e-3,
                               atol=1e-6)
    assert rel(q_sd2["embedding.2.weight"].grad[0], g["grad_embedding2_weight_row0"]) < 1e-3
