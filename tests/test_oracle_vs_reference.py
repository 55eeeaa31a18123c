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
