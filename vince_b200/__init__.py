"""vince_b200: B200-native (sm_100a) implementation of VINCE's encoder + InfoNCE + momentum-queue hot path.

Public surface mirrors the reference's own modules:

    vince_b200.VinceModel, vince_b200.VinceQueueModel     <- models/vince_model.py
    vince_b200.StorageQueue                               <- utils/storage_queue.py
    vince_b200.loss_util.similarity_cross_entropy         <- utils/loss_util.py
    vince_b200.backbone_models.ResNet18 / ResNet50        <- models/building_blocks/backbone_models.py
    vince_b200.optim.FusedSGD                             <- torch.optim.SGD as built at solvers/vince_solver.py:252-256
    vince_b200.knn.knn_eval / knn_accuracy                <- the kNN-CIFAR block of solvers/vince_solver.py:651-693

All arithmetic runs in libvince_b200.so (hand-written CUDA behind the C ABI of include/vince_b200.h).  There is no
CPU or PyTorch fallback: importing works anywhere, but every op raises on non-CUDA tensors or if the library is
not built (`python -m vince_b200.build`).
"""
from . import backbone_models, knn, loss_util, ops, optim  # noqa: F401
from .backbone_models import ResNet18, ResNet50  # noqa: F401
from .storage_queue import StorageQueue  # noqa: F401
from .vince_model import VinceModel, VinceQueueModel  # noqa: F401

__version__ = "0.2.0"
