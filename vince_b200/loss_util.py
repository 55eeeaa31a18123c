"""similarity_cross_entropy with the reference's literal signature (utils/loss_util.py:7-62), on the GPU.

The solver never calls this directly - `VinceModel.forward/loss/get_metrics` use the fused InfoNCE kernel that
never materialises the similarity matrix.  This explicit-matrix form exists for callers that already hold a
similarity matrix (and for parity tests against the reference function itself).

Semantics kept: temperature division, row max over ALL columns, each positive contrasted against the negatives
only, positives returned in column order, outputs `dists [n_feat, n_rows1, nP]`, `dist`, `softmax_weights`,
`softmax_weight`.  Only the equal-positives-per-row branch (loss_util.py:35-38, the one the hot path exercises,
SURVEY.md Appendix B.3) is implemented; unequal counts raise instead of silently switching to the float-mask
variant, and `n_rows1` must be 1 as at every reference call site (vince_model.py:270,278).
"""
import ctypes

import torch

from . import _lib, ops


def _run(similarities, temperature, mask):
    if hasattr(similarities, "materialize"):
        similarities = similarities.materialize()
    if not similarities.is_cuda:
        raise RuntimeError("vince_b200.loss_util: CUDA tensors only (no CPU fallback)")
    sims = similarities.detach().contiguous().float()
    R, C = sims.shape
    if mask.shape != sims.shape:
        raise AssertionError("mask.shape != similarities.shape")
    mask8 = mask.to(device=sims.device).contiguous().view(torch.uint8) if mask.dtype == torch.bool else \
        (mask != 0).contiguous().view(torch.uint8)
    counts = mask8.sum(-1)                       # argument validation (the reference syncs here too, :27-29)
    nP = int(counts[0].item())
    if not bool((counts == nP).all()):
        raise NotImplementedError("similarity_cross_entropy: rows with different numbers of positives (the "
                                  "reference's USE_FLOAT branch) are not implemented")
    dev = sims.device
    out = {k: torch.empty((R, nP), device=dev, dtype=torch.float32) for k in ("dists", "weights", "pos_sim")}
    out["neg_max"] = torch.empty((R,), device=dev, dtype=torch.float32)
    out["row_lse"] = torch.empty((R, 2), device=dev, dtype=torch.float32)
    out["scalars"] = torch.zeros((8,), device=dev, dtype=torch.float32)
    flag = torch.zeros((1,), device=dev, dtype=torch.int32)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().vince_masked_ce_fwd(p(sims), p(mask8), R, C, nP, float(temperature), p(out["dists"]),
                                                 p(out["weights"]), p(out["pos_sim"]), p(out["neg_max"]),
                                                 p(out["row_lse"]), p(out["scalars"]), p(flag), ops._stream()),
                   "vince_masked_ce_fwd")
    return out, nP


def similarity_cross_entropy(similarities, temperature, n_feat, n_rows1, mask=None, n_positives_per_row=None):
    if n_rows1 != 1:
        raise NotImplementedError("n_rows1 != 1 is never used by the reference's callers")
    if mask is None:
        assert n_positives_per_row is not None
        mask = torch.eye(n_feat, device=similarities.device, dtype=torch.bool).repeat_interleave(n_positives_per_row, 1)
    out, nP = _run(similarities, temperature, mask)
    return dict(
        dists=out["dists"].view(n_feat, n_rows1, nP),
        dist=out["scalars"][0],
        softmax_weights=out["weights"].view(n_feat, n_rows1, nP),
        softmax_weight=out["scalars"][1],
    )


def similarity_metrics(similarities, mask, temperature=1.0):
    """The metric quantities of VinceModel.get_metrics (vince_model.py:327-342) for an explicit matrix."""
    out, _ = _run(similarities, temperature, mask)
    sc = out["scalars"]
    return {"nce_accuracy": sc[2], "cosine_sim": sc[3], "cosine_sim_neg_max": sc[4]}
