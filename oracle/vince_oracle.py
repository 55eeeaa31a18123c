"""CPU oracle: a restatement of VINCE's encoder + InfoNCE + queue/EMA hot path.

*** TEST INFRASTRUCTURE.  NOT PART OF THE PRODUCT PATH. ***
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker / the CPU baseline.
vince_b200/ never imports it; the product path raises if the CUDA library is missing.

Parity pinning: the reference (danielgordon10/vince @ ae14e4ff) ships NO tests, golden
vectors or fixtures, so this oracle is pinned against outputs of the
reference's own code run in the dev container:
  * oracle/make_golden.py imports the unmodified reference (via oracle/ref_loader.py + the
    dg_util shim) and writes tests/golden/*.npz (committed, with the generating script);
  * tests/test_oracle_golden.py checks this file against those vectors everywhere;
  * tests/test_oracle_vs_reference.py checks it against the live reference where
    /root/reference is mounted.

Everything is written as plain functional torch on whatever dtype the inputs carry
(fp32 = the reference's arithmetic; fp64 = "truth" twin used to separate oracle rounding
from kernel rounding).  Weights travel as a flat dict using the reference's state_dict
key names (`feature_extractor.module.model.*`, `embedding.{0,2}.*`, `jigsaw_*`).

Reference anchors (file:line under /root/reference):
  models/building_blocks/backbone_models.py:39-54   Backbone.forward (children[0:8] = through layer4)
  models/building_blocks/resnet.py:76-92            BasicBlock.forward
  models/building_blocks/resnet.py:117-137          Bottleneck.forward
  models/building_blocks/resnet.py:170-179,231-247  stem + layers
  models/vince_model.py:123-196                     extract_features / get_embeddings
  models/vince_model.py:198-250                     forward (similarity matrices + masks)
  models/vince_model.py:252-349                     loss / get_metrics
  models/vince_model.py:587-595                     param_update (momentum EMA)
  utils/loss_util.py:7-62                           similarity_cross_entropy
  utils/storage_queue.py:4-56                       StorageQueue
  solvers/vince_solver.py:405-499                   run_train_iteration (the caller replayed by train_step)
"""
import collections
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # torch.nn.BatchNorm2d default, used by torchvision resnets
BN_MOMENTUM = 0.1

# (block kind, blocks per layer, output channels)   torchvision.models.resnet18/50 as wrapped by
# backbone_models.py:57-75
RESNET_SPECS = {
    "ResNet18": ("basic", (2, 2, 2, 2), 512),
    "ResNet50": ("bottleneck", (3, 4, 6, 3), 2048),
}
BACKBONE_PREFIX = "feature_extractor.module.model."


# --------------------------------------------------------------------------------------
# parameter construction (same distributions as torchvision/torch defaults; used where the
# reference itself is not importable, i.e. on the GPU box)
# --------------------------------------------------------------------------------------
def resnet_param_shapes(backbone):
    """Ordered (name, shape, kind) for the torchvision resnet incl. the unused `fc`
    (it is part of vince_parameters(), vince_model.py:96-104)."""
    kind, layers, _ = RESNET_SPECS[backbone]
    out = []

    def conv(name, cout, cin, k):
        out.append((name + ".weight", (cout, cin, k, k), "conv"))

    def bn(name, c):
        out.append((name + ".weight", (c,), "ones"))
        out.append((name + ".bias", (c,), "zeros"))
        out.append((name + ".running_mean", (c,), "zeros_buf"))
        out.append((name + ".running_var", (c,), "ones_buf"))
        out.append((name + ".num_batches_tracked", (), "count_buf"))

    conv("conv1", 64, 3, 7)
    bn("bn1", 64)
    inplanes = 64
    expansion = 1 if kind == "basic" else 4
    for li, nblocks in enumerate(layers):
        planes = 64 * (2 ** li)
        for b in range(nblocks):
            stride = 2 if (b == 0 and li > 0) else 1
            p = "layer%d.%d" % (li + 1, b)
            if kind == "basic":
                conv(p + ".conv1", planes, inplanes, 3)
                bn(p + ".bn1", planes)
                conv(p + ".conv2", planes, planes, 3)
                bn(p + ".bn2", planes)
            else:
                conv(p + ".conv1", planes, inplanes, 1)
                bn(p + ".bn1", planes)
                conv(p + ".conv2", planes, planes, 3)
                bn(p + ".bn2", planes)
                conv(p + ".conv3", planes * 4, planes, 1)
                bn(p + ".bn3", planes * 4)
            if b == 0 and (stride != 1 or inplanes != planes * expansion):
                conv(p + ".downsample.0", planes * expansion, inplanes, 1)
                bn(p + ".downsample.1", planes * expansion)
            inplanes = planes * expansion
    out.append(("fc.weight", (1000, 512 * expansion), "linear_w"))
    out.append(("fc.bias", (1000,), "linear_b:%d" % (512 * expansion)))
    return out


def make_state_dict(backbone="ResNet18", embedding_size=128, jigsaw=False, seed=0, dtype=torch.float32):
    """Random-init weights with torchvision's distributions (kaiming-normal fan_out convs, BN 1/0,
    nn.Linear default uniform).  Deterministic for a given torch version via a CPU generator."""
    g = torch.Generator().manual_seed(seed)
    sd = collections.OrderedDict()
    C = RESNET_SPECS[backbone][2]

    def linear(name, cout, cin):
        bound = 1.0 / math.sqrt(cin)
        sd[name + ".weight"] = (torch.rand((cout, cin), generator=g, dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)
        sd[name + ".bias"] = (torch.rand((cout,), generator=g, dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)

    for name, shape, kind in resnet_param_shapes(backbone):
        key = BACKBONE_PREFIX + name
        if kind == "conv":
            fan_out = shape[0] * shape[2] * shape[3]
            std = math.sqrt(2.0 / fan_out)
            sd[key] = (torch.randn(shape, generator=g, dtype=torch.float64) * std).to(dtype)
        elif kind in ("ones", "ones_buf"):
            sd[key] = torch.ones(shape, dtype=dtype)
        elif kind in ("zeros", "zeros_buf"):
            sd[key] = torch.zeros(shape, dtype=dtype)
        elif kind == "count_buf":
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif kind == "linear_w":
            bound = 1.0 / math.sqrt(shape[1])
            sd[key] = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)
        elif kind.startswith("linear_b"):
            bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
            sd[key] = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)
    linear("embedding.0", C, C)
    linear("embedding.2", embedding_size, C)
    if jigsaw:
        linear("jigsaw_linear", C, C)
        linear("jigsaw_embedding.0", C, 9 * C)
        linear("jigsaw_embedding.2", embedding_size, C)
    return sd


def vince_parameter_names(sd, jigsaw=False):
    """Names of the tensors VinceModel.vince_parameters() yields (vince_model.py:96-104):
    backbone *parameters* (incl. fc; NOT BN running stats) + embedding (+ jigsaw layers)."""
    names = [k for k in sd if k.startswith(BACKBONE_PREFIX)
             and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    names += [k for k in sd if k.startswith("embedding.")]
    if jigsaw:
        names += [k for k in sd if k.startswith("jigsaw_linear.")]
        names += [k for k in sd if k.startswith("jigsaw_embedding.")]
    return names


def clone_state_dict(sd, dtype=None):
    out = collections.OrderedDict()
    for k, v in sd.items():
        v = v.clone()
        if dtype is not None and v.is_floating_point():
            v = v.to(dtype)
        out[k] = v
    return out


# --------------------------------------------------------------------------------------
# ResNet forward (resnet.py:76-92, 117-137, 231-247), BN written out explicitly
# --------------------------------------------------------------------------------------
def batch_norm(x, sd, name, train):
    """nn.BatchNorm2d.forward: batch statistics (biased var) when train, running stats otherwise;
    running stats updated in place in `sd` with momentum 0.1 and the UNBIASED variance."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if train:
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        with torch.no_grad():
            rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
            rm.mul_(1 - BN_MOMENTUM).add_(mean.detach().to(rm.dtype), alpha=BN_MOMENTUM)
            rv.mul_(1 - BN_MOMENTUM).add_((var.detach() * (n / max(n - 1, 1))).to(rv.dtype), alpha=BN_MOMENTUM)
            if name + ".num_batches_tracked" in sd:
                sd[name + ".num_batches_tracked"] += 1
    else:
        mean, var = sd[name + ".running_mean"].to(x.dtype), sd[name + ".running_var"].to(x.dtype)
    scale = w / torch.sqrt(var + BN_EPS)
    shift = b - mean * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def _conv(x, sd, name, stride, padding):
    return F.conv2d(x, sd[name + ".weight"], None, stride=stride, padding=padding)


def _basic_block(x, sd, p, stride, train):
    identity = x
    out = F.relu(batch_norm(_conv(x, sd, p + ".conv1", stride, 1), sd, p + ".bn1", train))
    out = batch_norm(_conv(out, sd, p + ".conv2", 1, 1), sd, p + ".bn2", train)
    if (p + ".downsample.0.weight") in sd:
        identity = batch_norm(_conv(x, sd, p + ".downsample.0", stride, 0), sd, p + ".downsample.1", train)
    return F.relu(out + identity)


def _bottleneck(x, sd, p, stride, train):
    identity = x
    out = F.relu(batch_norm(_conv(x, sd, p + ".conv1", 1, 0), sd, p + ".bn1", train))
    out = F.relu(batch_norm(_conv(out, sd, p + ".conv2", stride, 1), sd, p + ".bn2", train))   # torchvision v1.5: stride on 3x3
    out = batch_norm(_conv(out, sd, p + ".conv3", 1, 0), sd, p + ".bn3", train)
    if (p + ".downsample.0.weight") in sd:
        identity = batch_norm(_conv(x, sd, p + ".downsample.0", stride, 0), sd, p + ".downsample.1", train)
    return F.relu(out + identity)


def resnet_forward(x, sd, backbone, train, prefix=BACKBONE_PREFIX, taps=None):
    """Backbone.forward with final_layer=-2: conv1,bn1,relu,maxpool,layer1..4 (backbone_models.py:39-54).
    NCHW in, NCHW `[B,C,h,w]` out.  `taps` (optional dict) receives intermediate activations."""
    kind, layers, _ = RESNET_SPECS[backbone]
    x = _conv(x, sd, prefix + "conv1", 2, 3)
    if taps is not None:
        taps["conv1_raw"] = x
    x = F.relu(batch_norm(x, sd, prefix + "bn1", train))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["maxpool"] = x
    block = _basic_block if kind == "basic" else _bottleneck
    for li, nblocks in enumerate(layers):
        for b in range(nblocks):
            stride = 2 if (b == 0 and li > 0) else 1
            x = block(x, sd, "%slayer%d.%d" % (prefix, li + 1, b), stride, train)
        if taps is not None:
            taps["layer%d" % (li + 1)] = x
    return x


def extract_features(x, sd, backbone, train, taps=None):
    """VinceModel.extract_features (vince_model.py:123-133) with the AdaptiveAvgPool2d branch."""
    spatial = resnet_forward(x, sd, backbone, train, taps=taps)
    return {"spatial_features": spatial, "extracted_features": spatial.mean(dim=(2, 3))}


def _linear(x, sd, name):
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


def jigsaw_patchify(data):
    """vince_model.py:144-155: pad to a multiple of 3 (bottom/right), [N,C,H,W] -> [9N,C,H/3,W/3], patches row-major."""
    if (data.shape[2] % 3) != 0 or (data.shape[3] % 3) != 0:
        data = F.pad(data, (0, 3 - data.shape[3] % 3, 0, 3 - data.shape[2] % 3))
    N, C, H, W = data.shape
    data = data.reshape(N, C, 3, H // 3, 3, W // 3).permute(0, 2, 4, 1, 3, 5).contiguous()
    return data.reshape(N * 9, C, H // 3, W // 3)


def get_embeddings(data, sd, backbone, train, shuffle_order=None, jigsaw=False, jigsaw_orders=None, taps=None):
    """VinceModel.get_embeddings (vince_model.py:135-196) with the random permutations INJECTED
    (`shuffle_order` replaces torch.randperm at :139; `jigsaw_orders [N,9]` replaces the per-row randperm(9) at :166).
    Returns the (un-split) dict: spatial_features, extracted_features, prenorm_features, embeddings."""
    if shuffle_order is not None:
        unshuffle = torch.zeros_like(shuffle_order)
        unshuffle[shuffle_order] = torch.arange(shuffle_order.numel())
        data = data[shuffle_order].contiguous()
    if jigsaw:
        data = jigsaw_patchify(data)
    out = extract_features(data, sd, backbone, train, taps=taps)
    feats = out["extracted_features"]
    if jigsaw:
        feats = _linear(feats, sd, "jigsaw_linear")
        feats = feats.reshape(-1, 9, feats.shape[1])
        rows = torch.arange(feats.shape[0])[:, None].expand(-1, 9)
        feats = feats[rows, jigsaw_orders].reshape(feats.shape[0], -1)
        feats = _linear(F.relu(_linear(feats, sd, "jigsaw_embedding.0")), sd, "jigsaw_embedding.2")
        out["extracted_features"] = feats            # overwritten with the [N,D] head output (:172)
        prenorm = feats
    else:
        prenorm = _linear(F.relu(_linear(feats, sd, "embedding.0")), sd, "embedding.2")
    out["prenorm_features"] = prenorm
    out["embeddings"] = F.normalize(prenorm, dim=1)          # eps 1e-12 (vince_model.py:180)
    if shuffle_order is not None:
        # :184-192 un-shuffles EVERY tensor with the N-long index (with jigsaw, spatial_features has 9N rows:
        # the reference's result for that entry is meaningless - we un-shuffle only N-row tensors)
        n = shuffle_order.numel()
        out = {k: (v[unshuffle] if v.shape[0] == n else v) for k, v in out.items()}
    return out


# --------------------------------------------------------------------------------------
# similarity matrices + masks (vince_model.py:50-77, 198-250)
# --------------------------------------------------------------------------------------
def block_diag_mask(n_rows, num_frames, n_cols_total):
    """similarity_mask / eye_mask sliced to the actual batch: [n_rows, n_cols_total] bool, block-diagonal
    nf x nf ones over the first n_rows columns (vince_model.py:52-77, :240)."""
    idx = torch.arange(n_rows) // max(num_frames, 1)
    m = torch.zeros((n_rows, n_cols_total), dtype=torch.bool)
    m[:, :n_rows] = idx[:, None] == idx[None, :]
    return m


def vince_forward(embeddings, queue_embeddings, queue_vectors, num_frames, inter_batch_comparison=True,
                  self_batch_comparison=False):
    """VinceModel.forward's similarity part.  Returns dict with vince_similarities(+_mask), vince_l_neg,
    optional vince_l_pos and vince_self_similarities(+_mask)."""
    out = {}
    B = embeddings.shape[0]
    if inter_batch_comparison:
        if self_batch_comparison:
            out["vince_self_similarities"] = embeddings @ embeddings.t()
            out["vince_self_similarities_mask"] = block_diag_mask(B, num_frames, B)
        negs = torch.cat((queue_embeddings, queue_vectors), dim=0)
        sims = embeddings @ negs.t()
        out["vince_l_neg"] = sims
        mask = block_diag_mask(B, num_frames, sims.shape[1])
    else:
        l_pos = (embeddings * queue_embeddings).sum(dim=1, keepdim=True)
        l_neg = embeddings @ queue_vectors.t()
        sims = torch.cat([l_pos, l_neg], dim=1)
        out["vince_l_pos"] = l_pos
        out["vince_l_neg"] = l_neg
        mask = torch.zeros(sims.shape, dtype=torch.bool)
        mask[:, 0] = True
    out["vince_similarities"] = sims
    out["vince_similarities_mask"] = mask
    return out


def similarity_cross_entropy(similarities, temperature, mask):
    """loss_util.similarity_cross_entropy with n_feat=B, n_rows1=1 and an equal number of positives per row
    (the branch the hot path exercises, loss_util.py:35-38).
    Returns dists [B,1,nP], dist, softmax_weights [B,1,nP], softmax_weight."""
    B = similarities.shape[0]
    z = similarities / temperature
    row_max = z.max(dim=-1, keepdim=True)[0]                       # over ALL columns (:24)
    s = z - row_max
    n_pos = int(mask[0].sum())
    assert bool((mask.sum(-1) == n_pos).all()), "oracle covers the equal-count branch only"
    neg = s[~mask].view(B, 1, -1)
    pos = s[mask].view(B, 1, n_pos)
    neg_exp = torch.exp(neg).sum(-1, keepdim=True)
    log_softmax = pos - torch.log(torch.exp(pos) + neg_exp)         # each positive vs negatives only (:40-43)
    dists = -log_softmax
    weights = torch.exp(log_softmax.detach())
    return {"dists": dists, "dist": dists.mean(), "softmax_weights": weights, "softmax_weight": weights.mean()}


def get_metrics(similarities, mask, softmax_weight, key=""):
    """VinceModel.get_metrics for one similarity matrix (vince_model.py:314-342)."""
    B = similarities.shape[0]
    pos_sim = similarities[mask].view(B, -1)
    neg_sim = similarities[~mask].view(B, -1)
    neg_max = neg_sim.max(dim=1, keepdim=True)[0]
    m = {
        "nce_accuracy_" + key + "mean": (pos_sim > neg_max).to(torch.float32).mean(),
        "nce_softmax_weight_" + key + "mean": softmax_weight,
        "cosine_" + key + "sim": pos_sim.mean(),
    }
    if key == "":
        m["cosine_sim_neg_max"] = neg_max.mean()
    return m


def infonce(embeddings, queue_embeddings, queue_vectors, num_frames, temperature, inter_batch_comparison=True,
            self_batch_comparison=False, self_temperature=0.03):
    """forward + loss + get_metrics in one call (vince_solver.py:424-426).  Returns (loss_dict, metrics, extras)."""
    fw = vince_forward(embeddings, queue_embeddings, queue_vectors, num_frames, inter_batch_comparison,
                       self_batch_comparison)
    ce = similarity_cross_entropy(fw["vince_similarities"], temperature, fw["vince_similarities_mask"])
    losses = {"nce_loss": ce["dist"]}
    metrics = get_metrics(fw["vince_similarities"], fw["vince_similarities_mask"], ce["softmax_weight"])
    extras = {"vince_loss_" + k: v for k, v in ce.items()}
    if inter_batch_comparison and self_batch_comparison:
        ce_s = similarity_cross_entropy(fw["vince_self_similarities"], self_temperature,
                                        fw["vince_self_similarities_mask"])
        losses["nce_loss_self"] = ce_s["dist"]
        metrics.update(get_metrics(fw["vince_self_similarities"], fw["vince_self_similarities_mask"],
                                   ce_s["softmax_weight"], key="self_"))
        extras.update({"vince_loss_self_" + k: v for k, v in ce_s.items()})
    extras.update(fw)
    return losses, metrics, extras


# --------------------------------------------------------------------------------------
# queue + EMA (storage_queue.py:4-56, vince_model.py:587-595)
# --------------------------------------------------------------------------------------
class StorageQueue:
    """Ring buffer [K,D]; only the vector part (the image / data-source lists are host bookkeeping)."""

    def __init__(self, maxsize, feat_size, init=None, dtype=torch.float32):
        self.maxsize, self.feat_size = maxsize, feat_size
        if init is None:
            init = F.normalize(torch.randn((maxsize, feat_size), dtype=dtype), dim=-1)
        self.vector_queue = init.clone()
        self.current_tail = 0
        self.full = False

    def enqueue(self, items):
        n = items.shape[0]
        if self.current_tail + n > self.maxsize:
            num_start = self.maxsize - self.current_tail
            if num_start > 0:
                self.vector_queue[self.current_tail:].copy_(items[:num_start])
            self.current_tail = 0
            self.full = True
            self.enqueue(items[num_start:])
        else:
            self.vector_queue[self.current_tail:self.current_tail + n].copy_(items)
            self.current_tail += n

    def dequeue(self):
        return {"queue_vectors": self.vector_queue}


def param_update(key_sd, query_sd, momentum, names):
    """theta_k <- m*theta_k + (1-m)*theta_q for every vince_parameters() tensor, in place on key_sd."""
    for n in names:
        key_sd[n].mul_(momentum).add_(query_sd[n], alpha=1 - momentum)


# --------------------------------------------------------------------------------------
# one scoring step = run_train_iteration:405-428 + :497-499 (no backward/SGD)
# --------------------------------------------------------------------------------------
def train_step(data, queue_data, query_sd, key_sd, queue, backbone, num_frames, temperature, momentum,
               shuffle_q=None, shuffle_k=None, inter_batch_comparison=True, self_batch_comparison=False,
               self_temperature=0.03, jigsaw=False, ema_names=None):
    """Key-encoder forward (no grad, train-mode BN) -> query-encoder forward -> InfoNCE + metrics against
    [keys || queue snapshot] -> enqueue keys -> EMA.  Mutates key_sd / query_sd (BN stats) and queue."""
    with torch.no_grad():
        k = get_embeddings(queue_data, key_sd, backbone, True, shuffle_order=shuffle_k)
    q = get_embeddings(data, query_sd, backbone, True, shuffle_order=shuffle_q)
    losses, metrics, extras = infonce(q["embeddings"], k["embeddings"], queue.vector_queue, num_frames, temperature,
                                      inter_batch_comparison, self_batch_comparison, self_temperature)
    queue.enqueue(k["embeddings"].detach())
    if ema_names is None:
        ema_names = vince_parameter_names(query_sd, jigsaw)
    with torch.no_grad():
        param_update(key_sd, query_sd, momentum, ema_names)
    return {"query": q, "key": k, "losses": losses, "metrics": metrics, "extras": extras}
