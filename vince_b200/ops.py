"""Tensor-level wrappers over the C ABI: validate (device / dtype / contiguity / shape) in Python, pass raw device
pointers + the current CUDA stream.  No op has a CPU or PyTorch fallback - non-CUDA tensors raise."""
import ctypes

import torch

from . import _lib

BN_MOMENTUM = 0.1
BN_EPS = 1e-5
PROFILE = None      # set to a list by bench.py to collect (name, algorithmic flops, start_event, end_event)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s: expected a torch.Tensor, got %r" % (name, type(t)))
    if not t.is_cuda:
        raise RuntimeError("%s: vince_b200 ops run on CUDA tensors only (got device %s); there is no CPU fallback"
                           % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s: expected dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s: tensor must be contiguous" % name)
    return ctypes.c_void_p(t.data_ptr())


def bn_side(raw, stats, bn):
    """bn: object with .weight .bias .running_mean .running_var .num_batches_tracked tensors.
    stats=None selects eval mode (running statistics)."""
    s = _lib.BnSide()
    s.raw = _ptr(raw, torch.float32, "bn.raw").value
    s.stats = _ptr(stats, torch.float64, "bn.stats").value if stats is not None else None
    s.gamma = _ptr(bn.weight, torch.float32, "bn.weight").value
    s.beta = _ptr(bn.bias, torch.float32, "bn.bias").value
    s.running_mean = _ptr(bn.running_mean, torch.float32, "bn.running_mean").value
    s.running_var = _ptr(bn.running_var, torch.float32, "bn.running_var").value
    nbt = getattr(bn, "num_batches_tracked", None)
    s.num_batches_tracked = _ptr(nbt, torch.int64, "bn.num_batches_tracked").value if nbt is not None else None
    return s


def conv_fwd(a_hi, a_lo, w_hi, w_lo, out, M, N, K, passes=3, geom=None, block_n=0, scale=None, bias=None, relu=False,
             stats=None):
    """geom=None: A is [M,K]; else dict(batch,H,W,Cin,R,S,stride,pad_lo_h,pad_lo_w,pad_hi_h,pad_hi_w) (NHWC gather)."""
    d = _lib.ConvDesc()
    d.a_hi = _ptr(a_hi, torch.bfloat16, "a_hi").value
    d.a_lo = _ptr(a_lo, torch.bfloat16, "a_lo").value if a_lo is not None else None
    d.w_hi = _ptr(w_hi, torch.bfloat16, "w_hi").value
    d.w_lo = _ptr(w_lo, torch.bfloat16, "w_lo").value if w_lo is not None else None
    d.out = _ptr(out, torch.float32, "out").value
    if out.numel() < M * N:
        raise ValueError("conv_fwd: output buffer too small")
    d.M, d.N, d.K = M, N, K
    if geom is not None:
        d.im2col = 1
        for k in ("batch", "H", "W", "Cin", "R", "S", "stride", "pad_lo_h", "pad_lo_w", "pad_hi_h", "pad_hi_w"):
            setattr(d, k, int(geom[k]))
    d.passes = passes
    d.block_n = block_n
    d.scale = _ptr(scale, torch.float32, "scale").value if scale is not None else None
    d.bias = _ptr(bias, torch.float32, "bias").value if bias is not None else None
    d.relu = 1 if relu else 0
    d.stats = _ptr(stats, torch.float64, "stats").value if stats is not None else None
    if PROFILE is not None:
        # bench.py's roofline leg: CUDA events around every tensor-core launch, on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().vince_conv_fwd(ctypes.byref(d), _stream()), "vince_conv_fwd")
        e1.record()
        k_true = geom.get("K_true", K) if geom is not None else K
        PROFILE.append(("conv_gemm", 2.0 * M * N * k_true, e0, e1))
        return
    _lib.check(_lib.lib().vince_conv_fwd(ctypes.byref(d), _stream()), "vince_conv_fwd")


def stem_geometry(H, W):
    """Shapes of the packed stem input and the equivalent 4x1 conv (see include/vince_b200.h: vince_stem_pack)."""
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    Hj = H // 2 + 1
    return dict(P=P, Q=Q, Hj=Hj,
                geom=dict(H=Hj, W=Q, Cin=64, R=4, S=1, stride=1, pad_lo_h=1, pad_lo_w=0, pad_hi_h=P + 2 - Hj,
                          pad_hi_w=0, K_true=147))      # algorithmic K of the 7x7x3 stem (the packed K is 256)


def stem_pack(x, gather_idx, x_hi, x_lo):
    N, C, H, W = x.shape
    if C != 3:
        raise ValueError("stem_pack: expected 3 input channels")
    _lib.check(_lib.lib().vince_stem_pack(_ptr(x, torch.float32, "x"), _ptr(gather_idx, torch.int64, "gather_idx"),
                                         _ptr(x_hi, torch.bfloat16, "x_hi"), _ptr(x_lo, torch.bfloat16, "x_lo"),
                                         N, H, W, _stream()), "vince_stem_pack")


def weight_prep(table_dev, n_entries, max_elems, w_hi, w_lo):
    _lib.check(_lib.lib().vince_weight_prep(_ptr(table_dev, torch.uint8, "table"), n_entries, max_elems,
                                           _ptr(w_hi, torch.bfloat16, "w_hi"), _ptr(w_lo, torch.bfloat16, "w_lo"),
                                           _stream()), "vince_weight_prep")


def bn_apply(main, M, C, relu, out_hi=None, out_lo=None, out_f32=None, res_planes=None, res_bn=None):
    res_kind, rh, rl, rb = 0, None, None, None
    if res_planes is not None:
        res_kind, rh, rl = 1, _ptr(res_planes[0], torch.bfloat16, "res_hi"), _ptr(res_planes[1], torch.bfloat16, "res_lo")
    elif res_bn is not None:
        res_kind, rb = 2, ctypes.byref(res_bn)
    _lib.check(_lib.lib().vince_bn_apply(ctypes.byref(main), res_kind, rh, rl, rb, 1 if relu else 0,
                                        _ptr(out_hi, torch.bfloat16, "out_hi"), _ptr(out_lo, torch.bfloat16, "out_lo"),
                                        _ptr(out_f32, torch.float32, "out_f32"), M, C, BN_MOMENTUM, BN_EPS, _stream()),
               "vince_bn_apply")


def bn_relu_maxpool(bn, out_hi, out_lo, N, P, Q, C):
    _lib.check(_lib.lib().vince_bn_relu_maxpool(ctypes.byref(bn), _ptr(out_hi, torch.bfloat16, "out_hi"),
                                               _ptr(out_lo, torch.bfloat16, "out_lo"), N, P, Q, C, BN_MOMENTUM, BN_EPS,
                                               _stream()), "vince_bn_relu_maxpool")


def bn_final_pool(main, N, HW, C, spatial_nchw, pooled, scatter_idx=None, res_planes=None, res_bn=None):
    res_kind, rh, rl, rb = 0, None, None, None
    if res_planes is not None:
        res_kind, rh, rl = 1, _ptr(res_planes[0], torch.bfloat16, "res_hi"), _ptr(res_planes[1], torch.bfloat16, "res_lo")
    elif res_bn is not None:
        res_kind, rb = 2, ctypes.byref(res_bn)
    _lib.check(_lib.lib().vince_bn_final_pool(ctypes.byref(main), res_kind, rh, rl, rb,
                                             _ptr(scatter_idx, torch.int64, "scatter_idx"),
                                             _ptr(spatial_nchw, torch.float32, "spatial"),
                                             _ptr(pooled, torch.float32, "pooled"), N, HW, C, BN_MOMENTUM, BN_EPS,
                                             _stream()), "vince_bn_final_pool")


def split_bf16(x, hi, lo):
    _lib.check(_lib.lib().vince_split_bf16(_ptr(x, torch.float32, "x"), _ptr(hi, torch.bfloat16, "hi"),
                                          _ptr(lo, torch.bfloat16, "lo"), x.numel(), _stream()), "vince_split_bf16")


def round_tf32(x, out):
    _lib.check(_lib.lib().vince_round_tf32(_ptr(x, torch.float32, "x"), _ptr(out, torch.float32, "out"), x.numel(),
                                          _stream()), "vince_round_tf32")


def l2_normalize(x, out, eps=1e-12):
    rows, D = x.shape
    _lib.check(_lib.lib().vince_l2_normalize(_ptr(x, torch.float32, "x"), _ptr(out, torch.float32, "out"), rows, D, eps,
                                            _stream()), "vince_l2_normalize")


def jigsaw_patchify(x, gather_idx, out):
    N, C, H, W = x.shape
    _lib.check(_lib.lib().vince_jigsaw_patchify(_ptr(x, torch.float32, "x"), _ptr(gather_idx, torch.int64, "idx"),
                                               _ptr(out, torch.float32, "out"), N, C, H, W, _stream()),
               "vince_jigsaw_patchify")


def jigsaw_gather(feats, order, out):
    N = order.shape[0]
    C = feats.shape[1]
    _lib.check(_lib.lib().vince_jigsaw_gather(_ptr(feats, torch.float32, "feats"), _ptr(order, torch.int64, "order"),
                                             _ptr(out, torch.float32, "out"), N, C, _stream()), "vince_jigsaw_gather")


def infonce_workspace_bytes(B, D):
    return int(_lib.lib().vince_infonce_workspace_bytes(B, D))


def infonce_fwd(q, keys, queue_tf32, num_frames, temperature, workspace=None):
    """Fused similarity + masked multi-positive cross entropy + metrics.  Returns a dict of device tensors:
    dists [B,nP], weights [B,nP], pos_sim [B,nP], neg_max [B], row_lse [B,2], scalars [8]."""
    B, D = q.shape
    K = 0 if queue_tf32 is None else queue_tf32.shape[0]
    nP = num_frames if num_frames > 0 else 1
    dev = q.device
    out = {
        "dists": torch.empty((B, nP), device=dev, dtype=torch.float32),
        "weights": torch.empty((B, nP), device=dev, dtype=torch.float32),
        "pos_sim": torch.empty((B, nP), device=dev, dtype=torch.float32),
        "neg_max": torch.empty((B,), device=dev, dtype=torch.float32),
        "row_lse": torch.empty((B, 2), device=dev, dtype=torch.float32),
        "scalars": torch.zeros((8,), device=dev, dtype=torch.float32),
    }
    if workspace is None:
        workspace = torch.empty((infonce_workspace_bytes(B, D),), device=dev, dtype=torch.uint8)
    d = _lib.InfoNceDesc()
    d.q = _ptr(q, torch.float32, "q").value
    d.keys = _ptr(keys, torch.float32, "keys").value
    d.queue_tf32 = _ptr(queue_tf32, torch.float32, "queue_tf32").value if K > 0 else None
    d.B, d.Bk, d.K, d.D, d.num_frames = B, keys.shape[0], K, D, num_frames
    d.temperature = float(temperature)
    for k in ("dists", "weights", "pos_sim", "neg_max", "row_lse", "scalars"):
        setattr(d, k, out[k].data_ptr())
    d.workspace = _ptr(workspace, torch.uint8, "workspace").value
    _lib.check(_lib.lib().vince_infonce_fwd(ctypes.byref(d), _stream()), "vince_infonce_fwd")
    out["_workspace"] = workspace     # keep alive until the stream has consumed it
    return out


def ema_enqueue(table_dev, n_chunks, momentum, queue=None, queue_tf32=None, keys=None, tail=0):
    """theta_k <- m theta_k + (1-m) theta_q over the chunk table and, if keys is given, ring-buffer enqueue at
    `tail` following storage_queue.py:31-49.  Returns (new_tail, wrapped)."""
    n0 = n1 = dst0 = dst1 = src1 = 0
    new_tail, wrapped = tail, False
    if keys is not None:
        K, D = queue.shape
        n = keys.shape[0]
        if n > K:
            raise ValueError("ema_enqueue: %d keys do not fit a queue of %d (enqueue in slices)" % (n, K))
        t = tail
        if t + n > K:
            first = K - t
            wrapped = True
            n0, dst0 = first * D, t * D
            n1, dst1, src1 = (n - first) * D, 0, first * D
            new_tail = n - first
        else:
            n0, dst0 = n * D, t * D
            new_tail = t + n
    _lib.check(_lib.lib().vince_ema_enqueue(
        _ptr(table_dev, torch.uint8, "ema table") if n_chunks else None, n_chunks, float(momentum),
        float(1 - momentum), _ptr(queue, torch.float32, "queue"), _ptr(queue_tf32, torch.float32, "queue_tf32"),
        _ptr(keys, torch.float32, "keys"), n0, dst0, n1, dst1, src1, _stream()), "vince_ema_enqueue")
    return new_tail, wrapped
