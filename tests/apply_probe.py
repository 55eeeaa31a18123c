"""Small ResNet-50 forwards that exercise every apply-epilogue route (train-mode two-pass on all expansions, eval-mode
folded BatchNorm) - for compute-sanitizer runs (GPU box only; not collected by pytest)."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vince_b200  # noqa: E402

dev = "cuda:0"
args = types.SimpleNamespace(
    backbone=vince_b200.ResNet50, num_frames=2, use_attention=False, feature_extractor_gpu_ids=[dev], pytorch_gpu_ids=[dev],
    vince_embedding_size=128, vince_queue_size=64, vince_temperature=0.07, vince_self_temperature=0.03,
    vince_momentum=0.999, jigsaw=False, inter_batch_comparison=True, self_batch_comparison=False, batch_size=4,
    use_imagenet=False)
torch.manual_seed(0)
model = vince_b200.VinceModel(args)
model.to(dev)
runner = model.feature_extractor.module.runner
runner.two_pass = 2
x = torch.randn(4, 3, 64, 64, device=dev)
outs = []
for train in (True, False):
    model.train(train)
    with torch.no_grad():
        for _ in range(2):
            e = model.get_embeddings({"data": x})["embeddings"]
    torch.cuda.synchronize()
    outs.append(e)
    print("train=%s embeddings finite=%s norm=%.4f" % (train, bool(torch.isfinite(e).all()), float(e.norm())))
