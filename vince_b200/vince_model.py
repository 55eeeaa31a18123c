"""VinceModel / VinceQueueModel on B200: the reference's API surface, hand-written sm_100a kernels underneath.

Drop-in for /root/reference/models/vince_model.py:19-349 (VinceModel) and :573-613 (VinceQueueModel) as used by
solvers/vince_solver.py:267-291,386-518: same constructor (`args`), attributes (`num_frames`, `output_channels`,
`feature_extractor.module`, `embedding`, `device`), methods and dict keys.  What changes is how the
work is done:

  get_embeddings   stem pack (+shuffle gather) -> tcgen05 implicit-GEMM convs with train-mode BN -> fused
                   relu(bn+residual)+avg-pool (+un-shuffle scatter) -> tcgen05 projection MLP -> L2 normalise
  forward/loss/    ONE fused InfoNCE kernel over [keys || queue]; the [B, B+K] similarity matrix is never written.
  get_metrics      `vince_similarities` / `vince_l_neg` in the returned dict are LazySimilarity handles that
                   materialise (through the same GEMM kernel) only if somebody asks.
  param_update     one multi-tensor EMA launch (optionally fused with the queue enqueue)

Forward only: autograd does not flow through the CUDA path (the query-encoder backward is SURVEY.md 8f rank 1,
not built yet); `loss()` returns detached tensors.
"""
import copy
import os
import time
import warnings
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import loss_util, ops
from .backbone_models import DataParallelShim
from .encoder import HeadRunner


def _device_of(t):
    return t if isinstance(t, torch.device) else torch.device(t)


# ---------------------------------------------------------------------------------------------------------------
# Encoder overlap.  The key-encoder forward (VinceQueueModel.__call__, vince_solver.py:405) and the query-encoder
# forward (VinceModel.get_embeddings, :406) are independent; each alternates tensor-core-bound convolutions with
# HBM-bound BatchNorm/ReLU passes, so running them on two CUDA streams lets one encoder's streaming kernels fill the
# memory pipe while the other's convolutions own the tensor cores.
#
# Safe by construction - nothing asynchronous ever leaks out of an API call: VinceQueueModel.forward only RECORDS a
# fork event on the caller's stream before launching the key encoder there (as always); the VinceModel.get_embeddings
# that IMMEDIATELY follows it (no other encoder-level call in between) on the same device, if it is handed the very
# batch tensor the fork saw (same storage, same version counter), runs the query encoder on a private side stream that waits for the fork event only, and the
# caller's stream waits for the side stream before get_embeddings returns.  Everything the caller launches afterwards
# is therefore ordered after BOTH encoders.  Any other call order simply runs on the caller's stream.
# Disable with args.vince_b200_overlap_encoders = False or VINCE_B200_OVERLAP=0.
# ---------------------------------------------------------------------------------------------------------------
_FORK = {}            # device index -> (event, data_ptr, version, call number) recorded by VinceQueueModel.forward
_SIDE_STREAMS = {}    # device index -> torch.cuda.Stream
_CALLS = [0]          # encoder-level API calls so far: a fork point is honoured only by the very next call


def _overlap_enabled(args):
    if os.environ.get("VINCE_B200_OVERLAP", "1") == "0":
        return False
    return bool(getattr(args, "vince_b200_overlap_encoders", True))


def _record_fork(inputs):
    data = inputs.get("data") if isinstance(inputs, dict) else None
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        return
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(data.device))
    _FORK[data.device.index] = (ev, data.data_ptr(), data._version, _CALLS[0])


def _take_fork(data):
    fork = _FORK.pop(data.device.index, None)
    if fork is None or fork[1] != data.data_ptr() or fork[2] != data._version or fork[3] != _CALLS[0] - 1:
        return None
    return fork[0]


def _side_stream(device):
    st = _SIDE_STREAMS.get(device.index)
    if st is None:
        st = _SIDE_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return st


def _record_stream_all(obj, stream):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream_all(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream_all(v, stream)


_PARAM_EPOCH = [0]      # bumped whenever a BaseModel may have re-allocated its parameters


TIME_STR = time.strftime("%Y_%m_%d_%H_%M_%S")     # constants.py:23 (one sub-directory per run, as in the reference)


def _checkpoint_files(folder):
    """All *.pt files under `folder` (recursively: the reference saves into checkpoint_dir/TIME_STR) as
    (iteration, mtime, path); the iteration is the trailing integer of the file name."""
    found = []
    for root, _, files in os.walk(folder):
        for f in files:
            if not f.endswith(".pt"):
                continue
            digits = "".join(ch if ch.isdigit() else " " for ch in os.path.splitext(f)[0]).split()
            it = int(digits[-1]) if digits else 0
            path = os.path.join(root, f)
            found.append((it, os.path.getmtime(path), path))
    return sorted(found)


def save_checkpoint(model, folder, num_to_keep, iteration):
    """dg_util pytorch_util.save semantics as used at models/base_model.py:23-25: write `folder/<iteration>.pt`
    (state_dict), then keep only the newest `num_to_keep` files of that folder (num_to_keep < 0 keeps all)."""
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, "%010d.pt" % iteration)
    torch.save(model.state_dict(), path)
    if num_to_keep > 0:
        files = sorted(f for f in os.listdir(folder) if f.endswith(".pt"))
        for f in files[:-num_to_keep]:
            os.remove(os.path.join(folder, f))
    return path


class BaseModel(nn.Module):
    """Stand-in for dg_util's pt_util.BaseModel + models/base_model.py:8-26 (device bookkeeping, save/restore)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self._device = "cpu"
        self.saves = 0

    @property
    def device(self):
        return self._device

    def to(self, device):
        self._device = device
        super().to(device)

    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float() ... may re-allocate parameters: cached device pointers (EMA chunk table) are stale
        _PARAM_EPOCH[0] += 1
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        _PARAM_EPOCH[0] += 1
        return super().load_state_dict(*a, **k)

    def restore(self, skip_filter=None) -> int:
        """models/base_model.py:13-19 -> dg_util restore_from_folder: load the newest checkpoint found under
        args.checkpoint_dir (searched recursively, so run directories written as checkpoint_dir/TIME_STR/ by `save`
        - ours or the reference's - are found), renaming keys that start with args.saved_variable_prefix to
        args.new_variable_prefix and dropping keys for which skip_filter(key) is true.  Returns the iteration encoded
        in the file name (0 when nothing was restored, with a warning)."""
        if not getattr(self.args, "restore", False):
            return 0
        ckpt_dir = self.args.checkpoint_dir
        files = _checkpoint_files(ckpt_dir) if os.path.isdir(ckpt_dir) else []
        if not files:
            warnings.warn("vince_b200 restore: no checkpoint (*.pt) found under %r; starting from iteration 0" % ckpt_dir)
            return 0
        iteration, _, path = files[-1]
        state = torch.load(path, map_location="cpu")
        if isinstance(state, dict) and "state_dict" in state and not any(torch.is_tensor(v) for v in state.values()):
            state = state["state_dict"]
        old = getattr(self.args, "saved_variable_prefix", None)
        new = getattr(self.args, "new_variable_prefix", None)
        if old is not None and new is not None:
            state = {(new + k[len(old):] if k.startswith(old) else k): v for k, v in state.items()}
        if skip_filter is not None:
            state = {k: v for k, v in state.items() if not skip_filter(k)}
        result = self.load_state_dict(state, strict=False)
        if result.missing_keys or result.unexpected_keys:
            warnings.warn("vince_b200 restore from %s: %d missing keys (e.g. %s), %d unexpected keys (e.g. %s)"
                          % (path, len(result.missing_keys), result.missing_keys[:3], len(result.unexpected_keys),
                             result.unexpected_keys[:3]))
        return iteration

    def save(self, iteration, num_to_keep=1):
        """models/base_model.py:21-26: rolling checkpoints under checkpoint_dir/TIME_STR, plus a never-pruned copy in
        long_save_checkpoint_dir every long_save_frequency saves."""
        if not getattr(self.args, "save", False):
            return
        save_checkpoint(self, os.path.join(self.args.checkpoint_dir, TIME_STR), num_to_keep, iteration)
        long_dir = getattr(self.args, "long_save_checkpoint_dir", None)
        freq = getattr(self.args, "long_save_frequency", 0)
        if long_dir and freq and self.saves > 0 and self.saves % freq == 0:
            save_checkpoint(self, long_dir, -1, iteration)
        self.saves += 1


class _FusedLoss(torch.autograd.Function):
    """Gives the fused kernel's loss scalar a grad_fn so that `loss.backward()` (vince_solver.py:463-469) reaches
    VinceModel._backward instead of dying inside autograd with "element 0 of tensors does not require grad"."""

    @staticmethod
    def forward(ctx, anchor, value, model, network_outputs, loss_name):
        ctx.model, ctx.network_outputs, ctx.loss_name = model, network_outputs, loss_name
        return value.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.model._backward(ctx.network_outputs, ctx.loss_name, grad_out)
        return None, None, None, None, None


class LazySimilarity:
    """Handle for a similarity matrix the fused kernel never wrote (vince_model.py:224,229-230 would).
    `.materialize()` computes q @ cols^T with the tcgen05 GEMM kernel and caches it."""

    def __init__(self, q, col_blocks, passes=3, prefix=None):
        self.q, self.col_blocks, self.passes = q, [c for c in col_blocks if c is not None and c.shape[0] > 0], passes
        self.prefix = prefix          # already-computed leading columns (MoCo's l_pos, vince_model.py:227-230)
        self.shape = (q.shape[0], sum(c.shape[0] for c in self.col_blocks) + (prefix.shape[1] if prefix is not None else 0))
        self._value = None

    def materialize(self):
        if self._value is None:
            B, D = self.q.shape
            dev = self.q.device
            Dp = (D + 63) // 64 * 64               # the GEMM wants K % 64 == 0: zero-pad the feature dim
            with torch.cuda.device(dev):
                def planes(x):
                    if Dp != D:
                        xp = torch.zeros((x.shape[0], Dp), device=dev, dtype=torch.float32)
                        xp[:, :D] = x
                        x = xp
                    hi = torch.empty(x.shape, device=dev, dtype=torch.float16)
                    lo = torch.empty_like(hi)
                    ops.split_f16(x.contiguous(), hi, lo)
                    return hi, lo
                q_hi, q_lo = planes(self.q)
                outs = []
                for cols in self.col_blocks:
                    n = cols.shape[0]
                    n_pad = (n + 31) // 32 * 32
                    c_hi, c_lo = planes(cols)
                    if n_pad != n:
                        pad = torch.zeros((n_pad - n, Dp), device=dev, dtype=torch.float16)
                        c_hi, c_lo = torch.cat((c_hi, pad)), torch.cat((c_lo, pad))
                    out = torch.empty((B, n_pad), device=dev, dtype=torch.float32)
                    ops.conv_fwd(q_hi, q_lo, c_hi, c_lo, out, B, n_pad, Dp, passes=3)
                    outs.append(out[:, :n])
                if self.prefix is not None:
                    outs.insert(0, self.prefix)
                self._value = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        return self._value

    def __repr__(self):
        return "LazySimilarity(shape=%s, materialized=%s)" % (self.shape, self._value is not None)


class VinceModel(BaseModel):
    def __init__(self, args):
        super(VinceModel, self).__init__(args)
        self.args = args
        self.num_frames = self.args.num_frames
        passes = getattr(args, "vince_b200_passes", 3)
        self.passes = passes

        self.feature_extractor = self.args.backbone(self.args, -2)
        resnet_output_channels = self.feature_extractor.output_channels
        self.output_channels = resnet_output_channels
        if getattr(self.args, "use_attention", False):
            raise NotImplementedError("--use-attention is not used by any reference config and is not implemented")
        # average_layers has no parameters in the reference either (AdaptiveAvgPool2d + RemoveDim, vince_model.py:33);
        # the pooling itself is fused into the last residual block's epilogue kernel.
        self.average_layers = nn.Sequential()
        self.feature_extractor = DataParallelShim(self.feature_extractor)
        self.feature_extractor_device = args.feature_extractor_gpu_ids[0]

        self.embedding = nn.Sequential(
            nn.Linear(self.output_channels, self.output_channels),
            nn.ReLU(inplace=True),
            nn.Linear(self.output_channels, self.args.vince_embedding_size),
        )
        self._heads = {"embedding": HeadRunner([self.embedding[0], self.embedding[2]], passes)}
        if self.args.jigsaw:
            self.jigsaw_linear = nn.Linear(self.output_channels, self.output_channels)
            self.jigsaw_embedding = nn.Sequential(
                nn.Linear(self.output_channels * 9, self.output_channels),
                nn.ReLU(inplace=True),
                nn.Linear(self.output_channels, self.args.vince_embedding_size),
            )
            self._heads["jigsaw"] = HeadRunner([self.jigsaw_linear, self.jigsaw_embedding[0], self.jigsaw_embedding[2]],
                                               passes)
        if getattr(self.args, "use_imagenet", False):
            raise NotImplementedError("--use-imagenet decoders are outside the hot path and are not implemented")
        self.launches = 0

    # the [B, B+K] boolean masks of vince_model.py:50-77 are only ever handed to the loss; build them on demand
    def _mask(self, n_rows, n_cols, num_frames, device):
        idx = torch.arange(n_rows, device=device) // max(num_frames, 1)
        m = torch.zeros((n_rows, n_cols), dtype=torch.bool, device=device)
        m[:, :n_rows] = idx[:, None] == idx[None, :]
        return m

    @property
    def similarity_mask(self):
        return self._mask(self.args.batch_size, self.args.batch_size + self.args.vince_queue_size, self.num_frames,
                          self.device)

    @property
    def eye_mask(self):
        return self._mask(self.args.batch_size, self.args.batch_size + self.args.vince_queue_size, 1, self.device)

    def to(self, device):
        super(VinceModel, self).to(device)
        self.feature_extractor.to(self.feature_extractor_device if str(self.feature_extractor_device) != "cpu" else device)

    def vince_parameters(self):
        params = (list(self.feature_extractor.parameters()) + list(self.embedding.parameters())
                  + list(self.average_layers.parameters()))
        if self.args.jigsaw:
            params += list(self.jigsaw_linear.parameters()) + list(self.jigsaw_embedding.parameters())
        return params

    @staticmethod
    def split_dict_by_type(batch_types, batch_sizes, dict_to_split):
        # vince_model.py:106-121, verbatim semantics (incl. the len(val) == len(batch_types) quirk)
        num_total = 0
        mini_batch_list = []
        assert "queue_vectors" not in dict_to_split
        for ind, (batch_type, batch_size) in enumerate(zip(batch_types, batch_sizes)):
            mini_batch = {
                key: (val[ind] if len(val) == len(batch_types) else val[num_total: num_total + batch_size])
                for key, val in dict_to_split.items()
            }
            mini_batch["batch_type"] = batch_type
            mini_batch.pop("batch_types", None)
            mini_batch_list.append(mini_batch)
            num_total += batch_size
        return mini_batch_list

    def extract_features(self, inputs, run_average_layer=True, gather_idx=None, scatter_idx=None, patch_grid=1,
                         tape=False):
        return_val = {}
        spatial, pooled = self.feature_extractor(inputs, gather_idx=gather_idx, scatter_idx=scatter_idx,
                                                 want_pooled=True, patch_grid=patch_grid, tape=tape)
        self.launches += self.feature_extractor.module.runner.launches
        return_val["spatial_features"] = spatial
        if run_average_layer:
            return_val["extracted_features"] = pooled
        return return_val

    def get_embeddings(self, inputs, jigsaw=False, shuffle=False, jigsaw_orders=None, _may_overlap=True):
        """vince_model.py:135-196.  `jigsaw_orders` ([N,9] int64, optional) replaces the per-row randperm(9) draw.
        Runs on a side stream, concurrently with a key-encoder forward launched just before on the caller's stream,
        when that is provably safe (see "Encoder overlap" at the top of this file)."""
        data = inputs["data"]
        if not data.is_cuda:
            raise RuntimeError("vince_b200.VinceModel: input batch must already be on the GPU (the solver's "
                               "prefetch thread does the H2D copy, vince_solver.py:352-355); no CPU fallback")
        if _may_overlap:
            _CALLS[0] += 1
        fork = _take_fork(data) if (_may_overlap and _overlap_enabled(self.args)) else None
        if fork is None:
            return self._get_embeddings(inputs, jigsaw, shuffle, jigsaw_orders)
        main = torch.cuda.current_stream(data.device)
        side = _side_stream(data.device)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            return_val = self._get_embeddings(inputs, jigsaw, shuffle, jigsaw_orders)
        _record_stream_all(return_val, main)      # allocated on the side stream, consumed (and freed) on the caller's
        main.wait_stream(side)
        return return_val

    def _get_embeddings(self, inputs, jigsaw, shuffle, jigsaw_orders):
        self.launches = 0
        cross = getattr(self, "cross_shuffle", None) if shuffle else None
        if cross is not None:
            return self._get_embeddings_cross_shuffled(inputs, jigsaw, jigsaw_orders, cross)
        # A backward may follow (vince_solver.py:463-469) when gradients are enabled on a training-mode model with
        # trainable parameters: the encoder then keeps a tape (no buffer recycling, saved BatchNorm statistics)
        taping = (torch.is_grad_enabled() and self.training and not jigsaw
                  and any(p.requires_grad for p in self.embedding.parameters()))
        self._train_tape = None
        with torch.no_grad():
            data = inputs["data"]
            n = data.shape[0]
            shuffle_order = None
            if shuffle:
                shuffle_order = torch.randperm(n, device=data.device)          # vince_model.py:139
            with torch.cuda.device(data.device):
                # uint8 HWC frames go straight to the stem packing kernel, which normalises them (encoder.py)
                if data.dtype not in (torch.float32, torch.uint8):
                    data = data.float()
                if jigsaw:
                    return_val = self._jigsaw_embeddings(data, shuffle_order, jigsaw_orders)
                else:
                    # shuffle gather folded into the stem's loads, un-shuffle into the last block's stores
                    return_val = self.extract_features(data, gather_idx=shuffle_order, scatter_idx=shuffle_order,
                                                       tape=taping)
                    head = self._heads["embedding"]
                    head.refresh()
                    hidden = head.linear(0, return_val["extracted_features"], relu=True)
                    output = head.linear(1, hidden, relu=False)
                    if taping:
                        self._train_tape = dict(pooled=return_val["extracted_features"], hidden=hidden, prenorm=output,
                                                shuffle_order=shuffle_order)
                    self.launches += head.launches
                    return_val["prenorm_features"] = output
                    emb = torch.empty_like(output)
                    ops.l2_normalize(output, emb)
                    self.launches += 1
                    return_val["embeddings"] = emb
        if "batch_types" in inputs:
            return_val = self.split_dict_by_type(inputs["batch_types"], inputs["batch_sizes"], return_val)
        return return_val

    def _get_embeddings_cross_shuffled(self, inputs, jigsaw, jigsaw_orders, cross):
        """shuffle=True with `self.cross_shuffle` (vince_b200.distributed.CrossGpuShuffle) set: the batch shuffle of
        vince_model.py:137-142 spans the GLOBAL batch of all ranks, as it does in the reference where nn.DataParallel
        splits the shuffled batch across the GPUs - frames travel to the rank that forwards them (one all-to-all),
        BatchNorm sees a random mix of clips, and every [B, ...] output travels back to its owner (:184-192)."""
        with torch.no_grad():
            data = inputs["data"]
            shuffled, ctx = cross.exchange(data)
            sub = {"data": shuffled}
            out = self._get_embeddings(sub, jigsaw, False, jigsaw_orders)
            n = data.shape[0]
            back = {}
            for key, val in out.items():
                if isinstance(val, torch.Tensor) and val.shape[0] == n:
                    back[key] = cross.restore(val, ctx)
                else:
                    back[key] = val
        if "batch_types" in inputs:
            back = self.split_dict_by_type(inputs["batch_types"], inputs["batch_sizes"], back)
        return back

    def _jigsaw_embeddings(self, data, shuffle_order, jigsaw_orders):
        # vince_model.py:144-173.  On one device the batch shuffle only permutes rows, so instead of gathering the
        # images we keep them in place and move the per-row patch permutations to their un-shuffled rows.
        # The patchify itself (pad to a multiple of 3 with the :145-146 quirk, cut, row-major patch order) is folded into
        # the stem packing (vince_stem_pack_grid): the [9N,3,H/3,W/3] patch tensor is never written.
        N = data.shape[0]
        dev = data.device
        return_val = self.extract_features(data, patch_grid=3)
        feats = return_val["extracted_features"]                                 # [9N, C]
        if jigsaw_orders is None:
            # vince_model.py:166 draws randperm(9) per row in a Python loop; argsort of iid uniforms is the same
            # distribution (uniform over the 9! orders, rows independent) in two launches instead of ~3N
            jigsaw_orders = torch.rand((N, 9), device=dev).argsort(dim=1)
        jigsaw_orders = jigsaw_orders.to(device=dev, dtype=torch.int64)
        if shuffle_order is not None:
            orders = torch.empty_like(jigsaw_orders)
            orders[shuffle_order] = jigsaw_orders          # row i of the shuffled batch is original row shuffle_order[i]
            jigsaw_orders = orders
        head = self._heads["jigsaw"]
        head.refresh()
        feats = head.linear(0, feats, relu=False)                                # jigsaw_linear
        gathered = torch.empty((N, 9 * feats.shape[1]), device=dev, dtype=torch.float32)
        ops.jigsaw_gather(feats, jigsaw_orders.contiguous(), gathered)
        hidden = head.linear(1, gathered, relu=True)
        output = head.linear(2, hidden, relu=False)
        self.launches += head.launches + 2
        return_val["extracted_features"] = output                                # overwritten as in :172
        return_val["prenorm_features"] = output
        emb = torch.empty_like(output)
        ops.l2_normalize(output, emb)
        self.launches += 1
        return_val["embeddings"] = emb
        return return_val

    # ------------------------------------------------------------------------------------------
    def forward(self, inputs: Dict[str, torch.Tensor]):
        """vince_model.py:198-250 (similarities) fused with loss_util.similarity_cross_entropy and get_metrics."""
        _CALLS[0] += 1
        return_val = copy.copy(inputs)
        if inputs.get("data_source") == "IN":
            raise NotImplementedError("ImageNet decoder branch (--use-imagenet) is outside the hot path")
        output = return_val["embeddings"]
        if "queue_embeddings" in inputs and "vince_similarities" not in inputs:
            queue_embeddings = inputs["queue_embeddings"]
            queue_vectors = inputs["queue_vectors"]
            queue_tf32 = inputs.get("queue_vectors_tf32")
            dev = output.device
            ibc = bool(self.args.inter_batch_comparison)
            nf = int(inputs["num_frames"]) if ibc else 0
            with torch.no_grad(), torch.cuda.device(dev):
                D = output.shape[1]
                if D % 32 != 0 or D > 128:
                    if D > 128:
                        raise NotImplementedError("fused InfoNCE supports embedding sizes up to 128, got %d" % D)
                    # odd embedding sizes (never used by the reference configs): zero-pad the feature dim,
                    # which leaves every dot product unchanged
                    Dp = (D + 31) // 32 * 32

                    def pad(x):
                        xp = torch.zeros((x.shape[0], Dp), device=dev, dtype=torch.float32)
                        xp[:, :D] = x
                        return xp
                    output_k, queue_embeddings_k, queue_vectors_k = pad(output), pad(queue_embeddings), pad(queue_vectors)
                    queue_tf32 = None
                else:
                    output_k, queue_embeddings_k, queue_vectors_k = output, queue_embeddings, queue_vectors
                output_k, queue_embeddings_k = output_k.contiguous(), queue_embeddings_k.contiguous()
                if queue_tf32 is None or queue_tf32.shape != queue_vectors_k.shape:
                    queue_tf32 = torch.empty_like(queue_vectors_k)
                    ops.round_tf32(queue_vectors_k.contiguous(), queue_tf32)
                    self.launches += 1
                fused = {"main": ops.infonce_fwd(output_k, queue_embeddings_k, queue_tf32, nf,
                                                self.args.vince_temperature)}
                fused["main"]["_operands"] = (output_k, queue_embeddings_k, queue_tf32, nf, self.args.vince_temperature)
                self.launches += 4 if ibc else 3
                if ibc and self.args.self_batch_comparison:
                    fused["self"] = ops.infonce_fwd(output_k, output_k, None, nf, self.args.vince_self_temperature)
                    fused["self"]["_operands"] = (output_k, output_k, None, nf, self.args.vince_self_temperature)
                    self.launches += 4
                    return_val["vince_self_similarities"] = LazySimilarity(output, [output])
                    return_val["vince_self_similarities_mask"] = None
            return_val["_vince_fused"] = fused
            if ibc:
                sims = LazySimilarity(output, [queue_embeddings, queue_vectors])
                return_val["vince_l_neg"] = sims
            else:
                return_val["vince_l_neg"] = LazySimilarity(output, [queue_vectors])
                return_val["vince_l_pos"] = fused["main"]["pos_sim"]
                sims = LazySimilarity(output, [queue_vectors], prefix=fused["main"]["pos_sim"])
            return_val["vince_similarities"] = sims
            return_val["vince_similarities_mask"] = None      # positives are structural (block-diagonal); see _mask()
        return return_val

    def loss(self, network_outputs: Optional[Dict]) -> Dict[str, Optional[Tuple[float, torch.Tensor]]]:
        if network_outputs is None:
            losses = {"nce_loss": None}
            if self.args.self_batch_comparison:
                losses["nce_loss_self"] = None
            return losses
        losses = {}
        if "_vince_fused" in network_outputs:
            fused = network_outputs["_vince_fused"]
            for key, name in (("main", ""), ("self", "self_")):
                if key not in fused:
                    continue
                f = fused[key]
                B, nP = f["dists"].shape
                network_outputs.update({
                    "vince_loss_" + name + "dists": f["dists"].view(B, 1, nP),
                    "vince_loss_" + name + "dist": f["scalars"][0],
                    "vince_loss_" + name + "softmax_weights": f["weights"].view(B, 1, nP),
                    "vince_loss_" + name + "softmax_weight": f["scalars"][1],
                })
                lname = "nce_loss" if key == "main" else "nce_loss_self"
                value = f["scalars"][0]
                if torch.is_grad_enabled() and self.training:
                    value = _FusedLoss.apply(self._grad_anchor(value.device), value, self, network_outputs, lname)
                losses[lname] = (1.0, value)
        elif "vince_similarities" in network_outputs:
            # caller supplied an explicit similarity matrix: generic masked cross entropy kernel
            similarities = network_outputs["vince_similarities"]
            batch_size = similarities.shape[0]
            mask = network_outputs["vince_similarities_mask"]
            sl = loss_util.similarity_cross_entropy(similarities, self.args.vince_temperature, batch_size, 1, mask)
            network_outputs.update({"vince_loss_" + key: val for key, val in sl.items()})
            losses["nce_loss"] = (1.0, sl["dist"])
            if self.args.self_batch_comparison:
                sl = loss_util.similarity_cross_entropy(network_outputs["vince_self_similarities"],
                                                        self.args.vince_self_temperature, batch_size, 1,
                                                        network_outputs["vince_self_similarities_mask"])
                network_outputs.update({"vince_loss_self_" + key: val for key, val in sl.items()})
                losses["nce_loss_self"] = (1.0, sl["dist"])
        return losses

    def _grad_anchor(self, device):
        # a leaf that requires grad, so autograd records _FusedLoss; NOT a Parameter / buffer (state_dict, SGD and the
        # EMA see exactly the reference's tensors)
        a = self.__dict__.get("_anchor")
        if a is None or a.device != device:
            a = torch.zeros((), device=device, requires_grad=True)
            self.__dict__["_anchor"] = a
        return a

    supports_backward = True

    def _backward(self, network_outputs, loss_name, grad_out):
        """Called by autograd when the solver runs loss.backward() (vince_solver.py:463-469): d loss / d embeddings with
        the fused InfoNCE backward kernel, then the projection head and the ResNet trunk from the tape of the last
        training-mode forward (vince_b200/backward.py).  Gradients are accumulated into `param.grad`."""
        from .backward import EncoderBackward, GradSlots, HeadBackward
        tape = getattr(self, "_train_tape", None)
        if tape is None:
            raise NotImplementedError(
                "vince_b200: loss.backward() needs the tape of a training-mode get_embeddings() call made with gradients "
                "enabled on this model (the jigsaw branch and the cross-GPU shuffle have no backward yet: SURVEY.md 8f)")
        key = "main" if loss_name == "nce_loss" else "self"
        weight = float(grad_out)
        with torch.no_grad():
            dq = self.embedding_gradients(network_outputs, {loss_name: weight}, only=key)
            dev = dq.device
            slots = self.__dict__.get("_grad_slots")
            if slots is None:
                slots = self.__dict__["_grad_slots"] = GradSlots(list(self.parameters()))
            slots.begin(dev)
            d_pooled = HeadBackward.run([self.embedding[0], self.embedding[2]], tape, dq.contiguous(), slots)
            if tape["shuffle_order"] is not None:
                d_pooled = d_pooled.index_select(0, tape["shuffle_order"])     # internal row i = frame shuffle_order[i]
            runner = self.feature_extractor.module.runner
            eb = self.__dict__.get("_encoder_backward")
            if eb is None or eb.runner is not runner:
                eb = self.__dict__["_encoder_backward"] = EncoderBackward(runner)
            eb.run(d_pooled, slots)
            slots.end()
            self.backward_launches = eb.launches + 12
            sync = getattr(self, "grad_sync", None)
            if sync is not None:
                sync(slots)

    def embedding_gradients(self, network_outputs: Dict, loss_weights: Optional[Dict[str, float]] = None, only=None):
        """d(sum of weighted losses)/d(embeddings) [B, D] with the fused backward kernel: the tensor autograd hands to
        the projection head when vince_solver.py:465 calls loss.backward() (keys / queue are detached,
        vince_model.py:598,610).  `loss_weights` maps "nce_loss" / "nce_loss_self" to their weights (default 1.0, as
        VinceModel.loss returns).  Autograd itself does not flow through the CUDA path (module docstring)."""
        fused = network_outputs["_vince_fused"]
        loss_weights = loss_weights or {}
        dq = None
        with torch.no_grad(), torch.cuda.device(network_outputs["embeddings"].device):
            for key, lname in (("main", "nce_loss"), ("self", "nce_loss_self")):
                if key not in fused or (only is not None and key != only):
                    continue
                q, keys, queue_tf32, nf, T = fused[key]["_operands"]
                dq = ops.infonce_bwd(q, keys, queue_tf32, nf, T, fused[key], grad_dist=loss_weights.get(lname, 1.0),
                                     symmetric=(key == "self"), dq=dq, accumulate=dq is not None)
        D = network_outputs["embeddings"].shape[1]
        return dq[:, :D] if dq.shape[1] != D else dq

    def get_metrics(self, network_outputs: Optional[Dict]) -> Dict[str, Optional[float]]:
        with torch.no_grad():
            metrics = {}
            if network_outputs is None:
                metrics.update({"nce_accuracy_mean": None, "nce_softmax_weight_mean": None, "cosine_sim": None,
                                "cosine_sim_neg_max": None})
                if self.args.self_batch_comparison:
                    metrics.update({"nce_accuracy_self_mean": None, "nce_softmax_weight_self_mean": None,
                                    "cosine_self_sim": None})
                return metrics
            if "_vince_fused" in network_outputs:
                fused = network_outputs["_vince_fused"]
                for key, name in (("main", ""), ("self", "self_")):
                    if key not in fused:
                        continue
                    sc = fused[key]["scalars"]
                    metrics["nce_accuracy_" + name + "mean"] = sc[2]
                    metrics["nce_softmax_weight_" + name + "mean"] = sc[1]
                    metrics["cosine_" + name + "sim"] = sc[3]
                    if key == "main":
                        metrics["cosine_sim_neg_max"] = sc[4]
            elif "vince_similarities" in network_outputs:
                for key in ["", "self_"]:
                    if "vince_" + key + "similarities" in network_outputs:
                        m = loss_util.similarity_metrics(network_outputs["vince_" + key + "similarities"],
                                                         network_outputs["vince_" + key + "similarities_mask"])
                        metrics["nce_accuracy_" + key + "mean"] = m["nce_accuracy"]
                        metrics["nce_softmax_weight_" + key + "mean"] = network_outputs["vince_loss_" + key + "softmax_weight"]
                        metrics["cosine_" + key + "sim"] = m["cosine_sim"]
                        if key == "":
                            metrics["cosine_sim_neg_max"] = m["cosine_sim_neg_max"]
            return metrics

    def get_image_output(self, network_outputs) -> Dict[str, np.ndarray]:
        # vince_model.py:351-571 builds tensorboard mosaics with cv2 on the CPU - visualisation, out of scope.
        return {}


class VinceQueueModel(BaseModel):
    def __init__(self, args, encoder: VinceModel):
        super(VinceQueueModel, self).__init__(args)
        self.queue_network = copy.deepcopy(encoder)
        self.vince_momentum = self.args.vince_momentum
        for param in self.queue_network.parameters():
            param.requires_grad = False
        self._ema_table = None
        self._ema_key = None
        self._ema_probe = None
        self.launches = 0

    def to(self, device):
        super(VinceQueueModel, self).to(device)
        self.queue_network.to(device)
        self._device = device

    def _table(self, encoder_model):
        import numpy as np
        # fast path (every training step): same encoder object and EVERY parameter of both encoders still lives where
        # the cached table says (re-allocating a single tensor must not leave the kernel writing through a stale pointer)
        # (a full scan costs ~0.3 us per tensor - 100 us per step for the 2 x 161 tensors of a ResNet-50, more than the
        #  EMA kernel itself: every step checks the module-re-allocation epoch (bumped by BaseModel._apply /
        #  load_state_dict(assign=True)) and a spread sample of 16 pointers; every 64th step all of them)
        probe = self._ema_probe
        if probe is not None and probe[0] is encoder_model and probe[3] == _PARAM_EPOCH[0]:
            self._ema_calls = getattr(self, "_ema_calls", 0) + 1
            check = probe[1] if self._ema_calls % 64 == 0 else probe[2]
            if all(p.data_ptr() == ptr for p, ptr in check):
                return self._ema_table
        dst = self.queue_network.vince_parameters()
        src = encoder_model.vince_parameters()
        key = tuple(p.data_ptr() for p in dst) + tuple(p.data_ptr() for p in src)
        every = [(p, p.data_ptr()) for p in dst + src]
        step = max(1, len(every) // 16)
        # the storages are kept alive with the table: if a tensor is swapped out behind the probe's back (p.data = ...), the
        # kernel keeps writing into memory that is still allocated until the next full check rebuilds the table
        self._ema_probe = (encoder_model, every, every[::step] + every[-1:], _PARAM_EPOCH[0],
                           [p.untyped_storage() for p in dst + src])
        if self._ema_key != key:
            chunks = []
            for d, s in zip(dst, src):
                if d.shape != s.shape or d.dtype != torch.float32 or s.dtype != torch.float32:
                    raise ValueError("param_update: parameter mismatch")
                if not (d.is_cuda and s.is_cuda and d.device == s.device):
                    raise RuntimeError("vince_b200 param_update: both encoders must live on the same CUDA device")
                if not (d.is_contiguous() and s.is_contiguous()):
                    raise ValueError("param_update: parameters must be contiguous")
                n = d.numel()
                for off in range(0, n, 8192):
                    chunks.append((d.data_ptr() + 4 * off, s.data_ptr() + 4 * off, min(8192, n - off)))
            arr = np.array(chunks, dtype=np.int64).reshape(-1, 3)
            self._ema_table = (torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(dst[0].device), len(chunks))
            self._ema_key = key
        return self._ema_table

    def param_update(self, encoder_model: VinceModel, momentum: float, enqueue=None, gather=None):
        """vince_model.py:587-592 as ONE launch.  `enqueue=(storage_queue, keys, images, data_source)` additionally
        performs StorageQueue.enqueue in the same launch (the reference does it just before, vince_solver.py:497);
        with `gather` (a vince_b200.distributed.KeyGather) the keys of every rank are all-gathered first and the
        rank-ordered rows are enqueued - NCCL all-gather + one kernel."""
        table, n = self._table(encoder_model)
        dev = table.device
        with torch.no_grad(), torch.cuda.device(dev):
            if enqueue is None:
                ops.ema_enqueue(table, n, momentum)
            elif gather is not None:
                queue, keys, images, source = enqueue
                gather.enqueue(queue, keys, None, source, ema=(table, n, momentum))
                self.launches = 2
                return
            else:
                queue, keys, images, source = enqueue
                if keys.shape[0] > queue.maxsize:
                    raise ValueError("fused enqueue: batch larger than the queue")
                if queue._shadow_is_stale():
                    queue._refresh_shadow()
                tail = queue.current_tail
                ops.ema_enqueue(table, n, momentum, queue.vector_queue, queue.vector_queue_tf32,
                                keys.detach().contiguous(), tail)
                queue.bookkeep(keys.shape[0], images, source)
        self.launches = 1

    def vince_update(self, encoder_model, enqueue=None, gather=None):
        self.param_update(encoder_model, self.vince_momentum, enqueue=enqueue, gather=gather)

    def forward(self, inputs, jigsaw=False, shuffle=True, jigsaw_orders=None):
        with torch.no_grad():
            queue_data = inputs["queue_data"]
            _CALLS[0] += 1
            if _overlap_enabled(self.args):
                _record_fork(inputs)          # lets the query encoder that follows run beside this one
            sub = {"data": queue_data}
            if "batch_types" in inputs:
                sub.update({"batch_types": inputs["batch_types"], "batch_sizes": inputs["batch_sizes"]})
            output_mini_batches = self.queue_network.get_embeddings(sub, jigsaw=jigsaw, shuffle=shuffle,
                                                                    jigsaw_orders=jigsaw_orders, _may_overlap=False)
            self.launches = self.queue_network.launches
            single = isinstance(output_mini_batches, dict)
            if single:
                output_mini_batches = [output_mini_batches]
            return_vals = []
            for outputs in output_mini_batches:
                return_val = {}
                for key, val in outputs.items():
                    if isinstance(val, torch.Tensor):
                        val = val.detach()
                    return_val["queue_" + key] = val
                return_vals.append(return_val)
            return return_vals
