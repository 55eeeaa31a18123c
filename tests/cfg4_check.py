"""Functional check of BASELINE.json configs[4] at full size on one GPU (GPU box only; not collected by pytest):
ResNet50 + jigsaw branch, 4 views, batch=128 frames (+ 9 patches of 75x75 each for the jigsaw encoder), K=131072.

    python tests/cfg4_check.py

Runs the solver's call sequence (vince_solver.py:397-428,497-499; the jigsaw coin flip fixed to "queue encoder gets the
patches"), checks every output is finite, and re-scores the GPU embeddings with the CPU oracle's InfoNCE."""
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vince_oracle as vo  # noqa: E402

import vince_b200  # noqa: E402


def main():
    dev = "cuda:0"
    B, nf, K, D, T = 128, 4, 131072, 128, 0.2
    args = types.SimpleNamespace(
        backbone=vince_b200.ResNet50, num_frames=nf, use_attention=False, feature_extractor_gpu_ids=[dev],
        pytorch_gpu_ids=[dev], vince_embedding_size=D, vince_queue_size=K, vince_temperature=T,
        vince_self_temperature=0.03, vince_momentum=0.999, jigsaw=True, inter_batch_comparison=True,
        self_batch_comparison=False, batch_size=B, use_imagenet=False)
    torch.manual_seed(0)
    model = vince_b200.VinceModel(args)
    model.to(dev)
    model.train()
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(dev)
    qm.train()
    queue = vince_b200.StorageQueue(K, D, device=dev)
    g = torch.Generator().manual_seed(3)
    batch = {"data": torch.randn((B, 3, 224, 224), generator=g).to(dev),
             "queue_data": torch.randn((B, 3, 224, 224), generator=g).to(dev),
             "batch_types": ["video"], "batch_sizes": [B], "data_source": "YT", "num_frames": nf}
    times = []
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        queue_batches = qm(batch, jigsaw=True, shuffle=True)          # :399-405 with the coin flip on the queue side
        outputs = model.get_embeddings(batch, jigsaw=False, shuffle=True)
        output = outputs[0]
        output.update(queue.dequeue())
        output.update({"data_source": "YT", "num_frames": nf})
        output.update(queue_batches[0])
        snapshot = queue.vector_queue.clone()
        output.update(model(output))
        loss = model.loss(output)["nce_loss"][1]
        metrics = model.get_metrics(output)
        qm.vince_update(model, enqueue=(queue, output["queue_embeddings"], [None] * B, "YT"))
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    for key in ("embeddings", "queue_embeddings", "extracted_features", "queue_prenorm_features"):
        assert torch.isfinite(output[key]).all(), key
    assert output["queue_embeddings"].shape == (B, D) and output["embeddings"].shape == (B, D)
    ref_losses, ref_metrics, _ = vo.infonce(output["embeddings"].cpu(), output["queue_embeddings"].cpu(), snapshot.cpu(),
                                            nf, T, True, False, 0.03)
    rel = abs(float(loss) - float(ref_losses["nce_loss"])) / abs(float(ref_losses["nce_loss"]))
    print("cfg4: loss %.6f (oracle on the same embeddings %.6f, rel %.2e)  acc %.3f  step %.1f ms (3rd iteration, "
          "R50 key encoder on 1152 patches of 75x75 + R50 query encoder on 128 frames of 224x224)"
          % (float(loss), float(ref_losses["nce_loss"]), rel, float(metrics["nce_accuracy_mean"]), times[-1] * 1e3))
    assert rel < 1e-3
    assert queue.current_tail == (3 * B) % K
    print("cfg4_check: PASS")


if __name__ == "__main__":
    main()
