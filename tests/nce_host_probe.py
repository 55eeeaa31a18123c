"""Host-side cost of the InfoNCE step through the reference-shaped API (GPU box only; not collected by pytest).

    python tests/nce_host_probe.py            # cProfile of 200 nce steps + wall / device time per step
"""
import cProfile
import os
import pstats
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vince_b200  # noqa: E402


def main():
    dev = "cuda:0"
    B, K, D, nf = 256, 65536, 128, 4
    backbone = sys.argv[1] if len(sys.argv) > 1 else "ResNet18"
    args = types.SimpleNamespace(
        backbone=getattr(vince_b200, backbone), num_frames=nf, use_attention=False, feature_extractor_gpu_ids=[dev],
        pytorch_gpu_ids=[dev], vince_embedding_size=D, vince_queue_size=K, vince_temperature=0.07,
        vince_self_temperature=0.03, vince_momentum=0.999, jigsaw=False, inter_batch_comparison=True,
        self_batch_comparison=False, batch_size=B, use_imagenet=False)
    torch.manual_seed(0)
    model = vince_b200.VinceModel(args)
    model.to(dev)
    model.train()
    qm = vince_b200.VinceQueueModel(args, model)
    qm.to(dev)
    queue = vince_b200.StorageQueue(K, D, device=dev)
    qv = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1)
    keys = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1)

    def nce_step():
        out = {"embeddings": qv, "extracted_features": qv, "queue_embeddings": keys, "data_source": "synthetic",
               "num_frames": nf}
        out.update(queue.dequeue())
        out.update(model(out))
        model.loss(out)
        model.get_metrics(out)
        qm.vince_update(model, enqueue=(queue, keys, [None] * B, "synthetic"), gather=None)

    with torch.no_grad():
        for _ in range(20):
            nce_step()
        torch.cuda.synchronize()
        for n in (200,):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(n):
                nce_step()
            e1.record()
            t_host = time.perf_counter() - t0
            torch.cuda.synchronize()
            print("%s: host issue time %.1f us/step, device time %.1f us/step" % (backbone, t_host / n * 1e6, e0.elapsed_time(e1) / n * 1e3))
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(200):
            nce_step()
        pr.disable()
        torch.cuda.synchronize()
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
