"""GPU k-nearest-neighbour evaluation of the encoder's embeddings.

Replaces the host-side kNN-CIFAR block of /root/reference/solvers/vince_solver.py:651-693 - embeddings of every image
through `model.get_embeddings` (eval-mode encoder, BatchNorm folded into the convolution epilogues), then sklearn
KDTree(all_features).query(k=11), drop the self match, scipy.stats.mode over the 10 neighbour labels,
mean(pred == label) - with one exact brute-force kernel on the device (`vince_knn_classify`).
"""
import torch

from . import _lib, ops


def knn_classify(features, labels, k=10):
    """features [n, D] fp32 CUDA, labels [n] int64 CUDA.  Returns (neighbour indices [n,k], distances [n,k], predicted
    label [n]) - neighbours exclude the nearest (self) match, as the reference's `neighbors[:, 1:]`."""
    n, D = features.shape
    features = features.contiguous()
    labels = labels.to(device=features.device, dtype=torch.int64).contiguous()
    nbr = torch.empty((n, k), device=features.device, dtype=torch.int64)
    dist = torch.empty((n, k), device=features.device, dtype=torch.float32)
    pred = torch.empty((n,), device=features.device, dtype=torch.int64)
    with torch.cuda.device(features.device):
        _lib.check(_lib.lib().vince_knn_classify(ops._ptr(features, torch.float32, "features"),
                                                 ops._ptr(labels, torch.int64, "labels"), n, D, k,
                                                 ops._ptr(nbr, torch.int64, "nbr_idx"), ops._ptr(dist, torch.float32, "nbr_dist"),
                                                 ops._ptr(pred, torch.int64, "pred"), ops._stream()), "vince_knn_classify")
    return nbr, dist, pred


def knn_accuracy(features, labels, k=10):
    """epoch_knn_cifar of vince_solver.py:679-680: mean(mode(labels[neighbours]) == labels), as a device scalar."""
    _, _, pred = knn_classify(features, labels, k)
    return (pred == labels.to(pred.device)).float().mean()


def knn_eval(model, images, labels, batch_size=256, k=10, mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD):
    """The whole CIFAR block of run_val: `images` is the dataset tensor the reference holds (`cifar_dataset.data`,
    [n,3,H,W], any real dtype, values in [0,255]-scale already divided as the dataset does) or raw uint8 HWC frames
    [n,H,W,3] (normalised on the fly by the stem kernel).  The model should be in eval mode (vince_solver.py:522)."""
    feats = []
    dev = model.device
    for s in range(0, images.shape[0], batch_size):
        data = images[s:s + batch_size].to(device=dev)
        if data.dtype != torch.uint8:
            m = torch.tensor(mean, device=dev, dtype=torch.float32).view(1, -1, 1, 1)
            sd = torch.tensor(std, device=dev, dtype=torch.float32).view(1, -1, 1, 1)
            data = (data.to(torch.float32) - m).div_(sd)                  # vince_solver.py:662-664
        feats.append(model.get_embeddings({"data": data})["embeddings"])
    feats = torch.cat(feats, dim=0)
    return knn_accuracy(feats, labels, k), feats
