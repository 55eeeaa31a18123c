import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_args(backbone="ResNet18", num_frames=4, batch_size=8, queue_size=1024, embedding_size=128, temperature=0.07,
              self_temperature=0.03, momentum=0.999, inter_batch_comparison=True, self_batch_comparison=False,
              jigsaw=False, device="cuda:0", passes=3):
    """The argparse fields the hot path reads, for vince_b200's classes."""
    import vince_b200
    return types.SimpleNamespace(
        backbone=getattr(vince_b200.backbone_models, backbone), num_frames=num_frames, use_attention=False,
        feature_extractor_gpu_ids=[device], pytorch_gpu_ids=[device], vince_embedding_size=embedding_size,
        vince_queue_size=queue_size, vince_temperature=temperature, vince_self_temperature=self_temperature,
        vince_momentum=momentum, jigsaw=jigsaw, inter_batch_comparison=inter_batch_comparison,
        self_batch_comparison=self_batch_comparison, batch_size=batch_size, use_imagenet=False,
        use_imagenet_weights=False, restore=False, save=False, checkpoint_dir="/tmp/_vince_b200_ckpt",
        vince_b200_passes=passes)


@pytest.fixture
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load
