"""GPU diagnostic (not collected by pytest): one training step of vince_b200 (forward, fused InfoNCE, loss.backward()
through the sm_100a backward kernels) against the oracle's autograd, printing the error of EVERY parameter gradient.

    python tests/train_probe.py [ResNet18|ResNet50] [B] [H] [fp64|fp32]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vince_oracle as vo  # noqa: E402
from conftest import make_args  # noqa: E402

import vince_b200  # noqa: E402

DEV = "cuda:0"


def oracle_grads(sd, data, queue_data, queue_init, backbone, nf, T, perm_q, perm_k, dtype=torch.float64):
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
    return losses["nce_loss"].detach(), {k: v.grad for k, v in q_sd.items() if v.is_floating_point() and v.grad is not None}, q_sd


def run_ours(backbone, B, H, nf, K, D, T, sd, data, queue_data, queue_init, perm_q, perm_k):
    from test_gpu_parity import injected_randperm
    args = make_args(backbone=backbone, num_frames=nf, batch_size=B, queue_size=K, embedding_size=D, temperature=T, device=DEV)
    model = vince_b200.VinceModel(args)
    model.load_state_dict(sd)
    model.to(DEV)
    model.train()
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(DEV)
    qm.train()
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
    return model, loss


def main():
    backbone = sys.argv[1] if len(sys.argv) > 1 else "ResNet18"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    dtype = torch.float32 if (len(sys.argv) > 4 and sys.argv[4] == "fp32") else torch.float64
    nf, K, D, T = 2, 256, 128, 0.07
    sd = vo.make_state_dict(backbone, D, seed=3)
    g = torch.Generator().manual_seed(17)
    data, queue_data = torch.randn((B, 3, H, H), generator=g), torch.randn((B, 3, H, H), generator=g)
    queue_init = F.normalize(torch.randn((K, D), generator=g), dim=-1)
    perm_q, perm_k = torch.randperm(B, generator=g), torch.randperm(B, generator=g)
    ref_loss, ref, _ = oracle_grads(sd, data, queue_data, queue_init, backbone, nf, T, perm_q, perm_k, dtype)
    if os.environ.get("PROBE_FP32_REF"):
        # how far is the reference's own fp32 autograd from the fp64 evaluation?
        _, ref32, _ = oracle_grads(sd, data, queue_data, queue_init, backbone, nf, T, perm_q, perm_k, torch.float32)
        num = sum(((ref32[k].double() - ref[k].double()).norm() ** 2).item() for k in ref)
        den = sum((ref[k].double().norm() ** 2).item() for k in ref)
        worst = max(((ref32[k].double() - ref[k].double()).norm() / ref[k].double().norm()).item() for k in ref)
        print("fp32 oracle autograd vs fp64: GLOBAL rel-L2 %.3e worst tensor %.3e" % ((num / den) ** 0.5, worst))
    model, loss = run_ours(backbone, B, H, nf, K, D, T, sd, data, queue_data, queue_init, perm_q, perm_k)
    print("loss %.6f oracle %.6f" % (loss.item(), ref_loss.item()))
    num = den = 0.0
    worst = 0.0
    for name, p in model.named_parameters():
        if name not in ref:
            print("%-60s (no oracle grad) ours=%s" % (name, None if p.grad is None else "set"))
            continue
        if p.grad is None:
            print("%-60s MISSING" % name)
            continue
        r = ref[name].double()
        e = (p.grad.detach().cpu().double() - r).norm().item()
        n = r.norm().item()
        num += e * e
        den += n * n
        rel = e / max(n, 1e-30)
        worst = max(worst, rel)
        if os.environ.get("PROBE_VERBOSE") or rel > 3e-3:
            print("%-60s rel %.3e  |ref| %.3e %s" % (name, rel, n, "<<<" if rel > 1e-2 else ""))
    print("GLOBAL rel-L2 %.3e  worst tensor %.3e" % ((num / den) ** 0.5, worst))


if __name__ == "__main__":
    main()
