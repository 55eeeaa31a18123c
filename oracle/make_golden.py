"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.  TEST INFRASTRUCTURE ONLY.

Run in the dev container (needs /root/reference):   python oracle/make_golden.py
The reference has no golden vectors of its own, so these files - outputs of the reference's
own VinceModel / VinceQueueModel / StorageQueue / loss_util code on seeded inputs - are what pins
oracle/vince_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path.

Weights come from vince_oracle.make_state_dict(seed) (a pure torch.Generator recipe, so the GPU box can
rebuild them bit-for-bit) and are loaded into the reference modules with load_state_dict; each file
records checksums of weights/inputs so RNG drift is detected instead of silently mis-compared.
Random permutations inside the reference (torch.randperm at vince_model.py:139,166) are injected by
patching torch.randperm for the duration of the call.
"""
import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import vince_oracle as vo  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def checksum(t):
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder(97.0) + 1.0
    return np.array([t.sum().item(), (t * w).sum().item(), t.abs().sum().item()])


def sd_checksum(sd):
    acc = np.zeros(3)
    for k, v in sd.items():
        if v.is_floating_point():
            acc += checksum(v)
    return acc


@contextlib.contextmanager
def injected_randperm(perms):
    """Make the next len(perms) torch.randperm calls return the given permutations."""
    real = torch.randperm
    it = iter(perms)

    def fake(n, *a, **k):
        p = next(it)
        assert p.numel() == n
        return p.clone()

    torch.randperm = fake
    try:
        yield
    finally:
        torch.randperm = real


def np_(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def build_reference_model(ref, backbone, nf, B, K, D, T, ibc=True, self_cmp=False, jigsaw=False, seed=0):
    args = ref_loader.make_args(backbone=backbone, num_frames=nf, batch_size=B, queue_size=K, embedding_size=D,
                                temperature=T, inter_batch_comparison=ibc, self_batch_comparison=self_cmp,
                                jigsaw=jigsaw)
    model = ref.VinceModel(args)
    sd = vo.make_state_dict(backbone, D, jigsaw=jigsaw, seed=seed)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    return args, model, sd


# ----------------------------------------------------------------------------------------------
def golden_infonce(ref):
    """loss_util.similarity_cross_entropy + VinceModel.forward/loss/get_metrics on small explicit inputs."""
    cases = [
        dict(name="ibc_nf4", B=8, K=64, D=16, nf=4, T=0.07, ibc=True, self_cmp=False),
        dict(name="ibc_nf2", B=8, K=40, D=16, nf=2, T=0.2, ibc=True, self_cmp=False),
        dict(name="ibc_nf1", B=6, K=32, D=8, nf=1, T=0.07, ibc=True, self_cmp=False),
        dict(name="ibc_self_nf4", B=8, K=64, D=16, nf=4, T=0.2, ibc=True, self_cmp=True),
        dict(name="moco", B=8, K=64, D=16, nf=4, T=0.07, ibc=False, self_cmp=False),
        dict(name="ibc_short_batch", B=8, K=64, D=16, nf=4, T=0.07, ibc=True, self_cmp=False, actual_B=4),
    ]
    out = {}
    g = torch.Generator().manual_seed(7)
    for c in cases:
        ref.loss_util.USE_FLOAT = None  # the module-global cache (loss_util.py:4) - reset per case
        args = ref_loader.make_args(backbone="ResNet18", num_frames=c["nf"], batch_size=c["B"], queue_size=c["K"],
                                    embedding_size=c["D"], temperature=c["T"], inter_batch_comparison=c["ibc"],
                                    self_batch_comparison=c["self_cmp"])
        # VinceModel.forward/loss/get_metrics never touch the network weights: call them unbound on a stub that
        # carries only args + masks, built by the reference's own __init__ code path for the masks.
        import types
        stub = types.SimpleNamespace(args=args, device="cpu", num_frames=c["nf"])
        # masks exactly as vince_model.py:50-77 builds them
        import scipy.linalg
        if c["ibc"]:
            if c["nf"] > 1:
                diag = torch.from_numpy(scipy.linalg.block_diag(
                    *[np.ones((c["nf"], c["nf"]), dtype=bool)] * (c["B"] // c["nf"])))
                stub.similarity_mask = torch.cat((diag, torch.zeros((c["B"], c["K"]), dtype=torch.bool)), dim=1)
            stub.eye_mask = torch.cat((torch.eye(c["B"], dtype=torch.bool),
                                       torch.zeros((c["B"], c["K"]), dtype=torch.bool)), dim=1)
        Bact = c.get("actual_B", c["B"])
        q = torch.nn.functional.normalize(torch.randn((Bact, c["D"]), generator=g), dim=1)
        k = torch.nn.functional.normalize(q + 0.5 * torch.randn((Bact, c["D"]), generator=g), dim=1)
        queue = torch.nn.functional.normalize(torch.randn((c["K"], c["D"]), generator=g), dim=1)
        inputs = {"extracted_features": q, "embeddings": q, "queue_embeddings": k, "queue_vectors": queue,
                  "data_source": "synthetic", "num_frames": c["nf"]}
        VM = ref.VinceModel
        fw = VM.forward(stub, inputs)
        losses = VM.loss(stub, fw)
        metrics = VM.get_metrics(stub, fw)
        rec = {"q": q, "k": k, "queue": queue,
               "cfg": np.array([c["B"], c["K"], c["D"], c["nf"], int(c["ibc"]), int(c["self_cmp"]), Bact]),
               "T": np.array([c["T"], args.vince_self_temperature]),
               "similarities": fw["vince_similarities"], "mask": fw["vince_similarities_mask"],
               "dists": fw["vince_loss_dists"], "dist": fw["vince_loss_dist"],
               "softmax_weights": fw["vince_loss_softmax_weights"], "softmax_weight": fw["vince_loss_softmax_weight"],
               "nce_loss": losses["nce_loss"][1]}
        for mk, mv in metrics.items():
            rec["metric_" + mk] = mv
        if c["self_cmp"]:
            rec["self_dists"] = fw["vince_loss_self_dists"]
            rec["nce_loss_self"] = losses["nce_loss_self"][1]
        # gradient of the loss wrt q (what the fused backward must reproduce), via the reference's own graph
        ref.loss_util.USE_FLOAT = None
        q2 = q.clone().requires_grad_(True)
        inputs2 = dict(inputs, embeddings=q2, extracted_features=q2)
        fw2 = VM.forward(stub, inputs2)
        l2 = VM.loss(stub, fw2)
        total = sum(v[0] * v[1] for v in l2.values())
        total.backward()
        rec["dq"] = q2.grad
        # oracle agreement at generation time
        o_losses, o_metrics, o_ex = vo.infonce(q, k, queue, c["nf"], c["T"], c["ibc"], c["self_cmp"],
                                               args.vince_self_temperature)
        assert torch.allclose(o_losses["nce_loss"], rec["nce_loss"], rtol=1e-6, atol=1e-6), c["name"]
        assert torch.equal(o_ex["vince_similarities_mask"], rec["mask"]), c["name"]
        for key, val in np_(rec).items():
            out["%s/%s" % (c["name"], key)] = val
    np.savez_compressed(os.path.join(OUT, "infonce.npz"), **out)
    print("infonce.npz:", len(out), "arrays")


def golden_queue(ref):
    """StorageQueue.enqueue wrap-around incl. recursion, tail==K edge, K not a multiple of n."""
    out = {}
    g = torch.Generator().manual_seed(11)
    cases = [
        dict(name="exact_multiple", K=32, D=8, sizes=[8, 8, 8, 8, 8]),
        dict(name="ragged", K=20, D=8, sizes=[8, 8, 8, 8, 3, 20]),
        dict(name="bigger_than_queue", K=10, D=4, sizes=[7, 25, 4]),
    ]
    for c in cases:
        q = ref.StorageQueue(c["K"], c["D"], device="cpu")
        init = q.vector_queue.clone()
        out[c["name"] + "/init"] = init.numpy()
        out[c["name"] + "/sizes"] = np.array(c["sizes"])
        oq = vo.StorageQueue(c["K"], c["D"], init=init)
        for i, n in enumerate(c["sizes"]):
            items = torch.randn((n, c["D"]), generator=g)
            q.enqueue(items, [None] * n, "src")
            oq.enqueue(items)
            out["%s/items%d" % (c["name"], i)] = items.numpy()
            out["%s/queue%d" % (c["name"], i)] = q.dequeue()["queue_vectors"].clone().numpy()
            out["%s/state%d" % (c["name"], i)] = np.array([q.current_tail, int(q.full)])
            assert torch.equal(oq.vector_queue, q.vector_queue) and oq.current_tail == q.current_tail and oq.full == q.full
    np.savez_compressed(os.path.join(OUT, "queue.npz"), **out)
    print("queue.npz:", len(out), "arrays")


def golden_encoder(ref):
    """VinceModel.get_embeddings (train + eval BN, with/without shuffle), R18 and R50, small images, plus
    VinceQueueModel.param_update on the same weights."""
    out = {}
    cases = [
        dict(name="r18_64", backbone="ResNet18", B=4, nf=2, H=64, W=64, D=128, seed=0, shuffle=True),
        dict(name="r50_64", backbone="ResNet50", B=4, nf=2, H=64, W=64, D=128, seed=1, shuffle=True),
        dict(name="r18_odd", backbone="ResNet18", B=3, nf=1, H=75, W=51, D=64, seed=2, shuffle=False),
    ]
    for c in cases:
        args, model, sd = build_reference_model(ref, c["backbone"], c["nf"], c["B"], 64, c["D"], 0.07, seed=c["seed"])
        g = torch.Generator().manual_seed(100 + c["seed"])
        x = torch.randn((c["B"], 3, c["H"], c["W"]), generator=g)
        perm = torch.randperm(c["B"], generator=g) if c["shuffle"] else None
        inputs = {"data": x, "batch_types": ["images"], "batch_sizes": [c["B"]]}
        with torch.no_grad():
            if perm is not None:
                with injected_randperm([perm]):
                    r = model.get_embeddings(inputs, shuffle=True)[0]
            else:
                r = model.get_embeddings(inputs, shuffle=False)[0]
        post = model.state_dict()
        rec = {"x": x, "perm": perm if perm is not None else torch.zeros(0, dtype=torch.int64),
               "cfg": np.array([c["B"], c["nf"], c["H"], c["W"], c["D"], c["seed"], int(c["shuffle"])]),
               "weights_checksum": sd_checksum(sd),
               "spatial_features": r["spatial_features"], "extracted_features": r["extracted_features"],
               "prenorm_features": r["prenorm_features"], "embeddings": r["embeddings"],
               "bn1_running_mean": post["feature_extractor.module.model.bn1.running_mean"],
               "bn1_running_var": post["feature_extractor.module.model.bn1.running_var"],
               "last_bn_running_var": post[[k for k in post if k.endswith("running_var")][-1]],
               "bn1_num_batches": post["feature_extractor.module.model.bn1.num_batches_tracked"]}
        # eval-mode forward with the updated running stats
        model.eval()
        with torch.no_grad():
            r_eval = model.get_embeddings({"data": x})
        rec["eval_embeddings"] = r_eval["embeddings"]
        rec["eval_extracted_features"] = r_eval["extracted_features"]
        model.train()
        # oracle agreement
        osd = vo.clone_state_dict(sd)
        o = vo.get_embeddings(x, osd, c["backbone"], True, shuffle_order=perm)
        err = (o["embeddings"] - r["embeddings"]).norm() / r["embeddings"].norm()
        print("  %s: oracle-vs-reference embedding rel-L2 = %.2e" % (c["name"], err.item()))
        assert err < 1e-4
        assert torch.allclose(osd["feature_extractor.module.model.bn1.running_var"], rec["bn1_running_var"], rtol=1e-5, atol=1e-6)
        # EMA: reference VinceQueueModel.param_update of a perturbed query onto the key copy
        qm = ref.VinceQueueModel(args, model)
        with torch.no_grad():
            for i, p in enumerate(model.vince_parameters()):
                p.add_(0.01 * ((i % 7) - 3))
        qm.param_update(model, 0.999)
        key_params = qm.queue_network.vince_parameters()
        rec["ema_checksum"] = sum(checksum(p) for p in key_params)
        rec["ema_embedding2_bias"] = dict(qm.queue_network.named_parameters())["embedding.2.bias"]
        rec["ema_n_tensors"] = np.array([len(key_params), sum(p.numel() for p in key_params)])
        for key, val in np_(rec).items():
            out["%s/%s" % (c["name"], key)] = val
    np.savez_compressed(os.path.join(OUT, "encoder.npz"), **out)
    print("encoder.npz:", len(out), "arrays")


def golden_jigsaw(ref):
    out = {}
    c = dict(name="r18_jigsaw", backbone="ResNet18", B=2, nf=2, H=50, W=50, D=32, seed=3)
    args, model, sd = build_reference_model(ref, c["backbone"], c["nf"], c["B"], 64, c["D"], 0.07, jigsaw=True,
                                            seed=c["seed"])
    g = torch.Generator().manual_seed(100 + c["seed"])
    x = torch.randn((c["B"], 3, c["H"], c["W"]), generator=g)
    perm = torch.randperm(c["B"], generator=g)
    orders = torch.stack([torch.randperm(9, generator=g) for _ in range(c["B"])])
    with torch.no_grad(), injected_randperm([perm] + list(orders)):
        r = model.get_embeddings({"data": x, "batch_types": ["images"], "batch_sizes": [c["B"]]}, jigsaw=True,
                                 shuffle=True)[0]
    rec = {"x": x, "perm": perm, "orders": orders, "weights_checksum": sd_checksum(sd),
           "cfg": np.array([c["B"], c["nf"], c["H"], c["W"], c["D"], c["seed"]]),
           "embeddings": r["embeddings"], "prenorm_features": r["prenorm_features"]}
    o = vo.get_embeddings(x, vo.clone_state_dict(sd), c["backbone"], True, shuffle_order=perm, jigsaw=True,
                          jigsaw_orders=orders)
    err = (o["embeddings"] - r["embeddings"]).norm() / r["embeddings"].norm()
    print("  jigsaw: oracle-vs-reference embedding rel-L2 = %.2e" % err.item())
    assert err < 1e-4
    for key, val in np_(rec).items():
        out["%s/%s" % (c["name"], key)] = val
    np.savez_compressed(os.path.join(OUT, "jigsaw.npz"), **out)
    print("jigsaw.npz:", len(out), "arrays")


def golden_step_cfg0(ref):
    """BASELINE.json configs[0]: R18, 2 views/clip, batch 8, K=1024, D=128, 224x224 - one full scoring step
    replaying vince_solver.py:405-428,497-499 with the reference's classes.  Inputs are regenerated from
    seeds at test time (too big to commit); checksums guard the regeneration."""
    B, nf, K, D, T, m = 8, 2, 1024, 128, 0.07, 0.999
    ref.loss_util.USE_FLOAT = None
    args, model, sd = build_reference_model(ref, "ResNet18", nf, B, K, D, T, seed=0)
    qm = ref.VinceQueueModel(args, model)
    qm.train()
    g = torch.Generator().manual_seed(1234)
    data = torch.randn((B, 3, 224, 224), generator=g)
    queue_data = torch.randn((B, 3, 224, 224), generator=g)
    queue_init = torch.nn.functional.normalize(torch.randn((K, D), generator=g), dim=-1)
    perm_k = torch.randperm(B, generator=g)
    perm_q = torch.randperm(B, generator=g)
    queue = ref.StorageQueue(K, D, device="cpu")
    queue.vector_queue = queue_init.clone()
    queue.current_tail = K - 3          # force a wrap inside this step
    batch = {"data": data, "queue_data": queue_data, "batch_types": ["images"], "batch_sizes": [B],
             "data_source": "synthetic", "num_frames": nf}
    with injected_randperm([perm_k, perm_q]):
        queue_batches = qm(batch, shuffle=True)
        outputs = model.get_embeddings(batch, shuffle=True)
    output = outputs[0]
    output.update(queue.dequeue())
    output.update({"data_source": "synthetic", "num_frames": nf})
    output.update(queue_batches[0])
    output.update(model(output))
    loss = model.loss(output)
    metrics = model.get_metrics(output)
    total = sum(v[0] * v[1] for v in loss.values())
    model.zero_grad()
    total.backward()
    grad_emb2 = model.embedding[2].weight.grad.clone()
    grad_conv1 = model.feature_extractor.module.model.conv1.weight.grad.clone()
    queue.enqueue(output["queue_embeddings"], [None] * B, "synthetic")
    qm.vince_update(model)
    key_params = qm.queue_network.vince_parameters()
    rec = {"cfg": np.array([B, nf, K, D]), "T_m": np.array([T, m]),
           "weights_checksum": sd_checksum(sd), "data_checksum": checksum(data),
           "queue_data_checksum": checksum(queue_data), "queue_init_checksum": checksum(queue_init),
           "perm_k": perm_k, "perm_q": perm_q,
           "embeddings": output["embeddings"], "queue_embeddings": output["queue_embeddings"],
           "extracted_features": output["extracted_features"],
           "loss": loss["nce_loss"][1], "dists": output["vince_loss_dists"],
           "queue_tail_rows": queue.vector_queue[K - 3:].clone(), "queue_head_rows": queue.vector_queue[:B].clone(),
           "queue_state": np.array([queue.current_tail, int(queue.full)]),
           "queue_checksum": checksum(queue.vector_queue),
           "ema_checksum": sum(checksum(p) for p in key_params),
           "grad_embedding2_weight_checksum": checksum(grad_emb2), "grad_conv1_checksum": checksum(grad_conv1),
           "grad_embedding2_weight_row0": grad_emb2[0].clone()}
    for mk, mv in metrics.items():
        rec["metric_" + mk] = mv
    # oracle agreement
    q_sd, k_sd = vo.clone_state_dict(sd), vo.clone_state_dict(sd)
    oq = vo.StorageQueue(K, D, init=queue_init)
    oq.current_tail = K - 3
    o = vo.train_step(data, queue_data, q_sd, k_sd, oq, "ResNet18", nf, T, m, shuffle_q=perm_q, shuffle_k=perm_k)
    print("  cfg0: ref loss %.6f oracle loss %.6f" % (rec["loss"].item(), o["losses"]["nce_loss"].item()))
    assert abs(rec["loss"].item() - o["losses"]["nce_loss"].item()) < 1e-4
    assert torch.allclose(oq.vector_queue, queue.vector_queue, atol=1e-5)
    np.savez_compressed(os.path.join(OUT, "step_cfg0.npz"), **np_(rec))
    print("step_cfg0.npz written")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = ref_loader.load_reference()
    golden_infonce(ref)
    golden_queue(ref)
    golden_encoder(ref)
    golden_jigsaw(ref)
    golden_step_cfg0(ref)


if __name__ == "__main__":
    main()
