"""Tensor-level wrappers over the C ABI: validate (device / dtype / contiguity / shape) in Python, pass raw device
pointers + the current CUDA stream.  No op has a CPU or PyTorch fallback - non-CUDA tensors raise.

Every op exists in two forms: `build_<op>(...)` validates once and returns a zero-argument callable with all ctypes
arguments pre-marshalled (the encoder plan replays lists of these, so the per-launch host cost is one ctypes call),
and `<op>(...)` = build + run for one-off use.
"""
import ctypes

import torch

from . import _lib

BN_MOMENTUM = 0.1
BN_EPS = 1e-5
PROFILE = None      # set to a list by bench.py to collect (name, algorithmic flops, start_event, end_event)


class GraphReplay:
    """A static launch list (pre-marshalled closures that all launch on torch's current stream, touch only buffers
    that stay allocated, and make no allocation or synchronisation) replayed as ONE CUDA graph launch: the first call
    runs the list eagerly (module loading, cudaFuncSetAttribute and other one-time work must not happen under stream
    capture), the second captures it, later calls replay the graph.  Per-launch host cost drops from one ctypes call
    per kernel to one cudaGraphLaunch per list, and the device sees no launch gaps.
    VINCE_B200_GRAPH=0 (or a bench.py PROFILE run, which wants events around single launches) keeps the eager path."""
    enabled = None

    def __init__(self, launches, pre=None):
        self.launches = launches
        self.pre = pre                  # optional callable run (and captured) before the launches, e.g. a memset
        self.graph = None
        self.calls = 0
        if GraphReplay.enabled is None:
            import os
            GraphReplay.enabled = os.environ.get("VINCE_B200_GRAPH", "1") != "0"

    def _eager(self):
        if self.pre is not None:
            self.pre()
        for run in self.launches:
            run()

    def __call__(self):
        if not GraphReplay.enabled or PROFILE is not None or not self.launches:
            return self._eager()
        if self.graph is not None:
            return self.graph.replay()
        self.calls += 1
        if self.calls < 2 or torch.cuda.is_current_stream_capturing():
            return self._eager()
        cur = torch.cuda.current_stream()
        g = torch.cuda.CUDAGraph()
        # capture on a side stream ordered after the caller's stream; capture executes nothing.  (Not the
        # `with torch.cuda.graph(...)` helper: it synchronises the device and empties the caching allocator first, which
        # cost a ResNet-50 training step hundreds of milliseconds of re-allocation when a capture fell into a timed region.)
        side = torch.cuda.Stream(device=cur.device)
        side.wait_stream(cur)
        ok = True
        with torch.cuda.stream(side):
            try:
                g.capture_begin(capture_error_mode="thread_local")
                try:
                    self._eager()
                finally:
                    g.capture_end()
            except Exception:
                ok = False
        cur.wait_stream(side)
        if not ok:
            GraphReplay.enabled = False      # e.g. a driver without the needed capture support: stay eager, loudly once
            import warnings
            warnings.warn("vince_b200: CUDA graph capture failed, replaying launch lists eagerly")
            torch.cuda.synchronize()
            return self._eager()
        self.graph = g
        g.replay()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_handle():
    """cudaStream_t of torch's current stream on the current device, as an int.  torch.cuda.current_stream() builds a
    Stream object through three layers of device-index helpers (~8 us per call, measured - more than a kernel launch);
    the raw getter underneath it costs ~0.3 us."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _stream():
    return ctypes.c_void_p(_stream_handle())


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


def _val(t, dtype, name):
    p = _ptr(t, dtype, name)
    return p.value if p is not None else None


def _bind(fn, what, *args):
    """Callable that launches fn(*args, current_stream) and raises on a non-zero status."""
    check = _lib.check

    def run():
        check(fn(*args, ctypes.c_void_p(_stream_handle())), what)
    return run


def bn_side(raw, coef):
    """raw: [M,C] fp32 conv output; coef: [2*C] fp32 (scale, shift) from the conv's fused BN finalize or bn_eval_coef."""
    s = _lib.BnSide()
    s.raw = _val(raw, torch.float32, "bn.raw")
    s.coef = _val(coef, torch.float32, "bn.coef")
    s._keep = (raw, coef)
    return s


def build_bn_eval_coef(bn, coef):
    """eval-mode BatchNorm: coef <- (gamma / sqrt(running_var + eps), beta - running_mean * scale)"""
    C = bn.weight.numel()
    run = _bind(_lib.lib().vince_bn_eval_coef, "vince_bn_eval_coef", _ptr(bn.weight, torch.float32, "bn.weight"),
                _ptr(bn.bias, torch.float32, "bn.bias"), _ptr(bn.running_mean, torch.float32, "bn.running_mean"),
                _ptr(bn.running_var, torch.float32, "bn.running_var"), BN_EPS, _ptr(coef, torch.float32, "coef"), C)
    run._keep = (bn, coef)
    return run


# ---------------------------------------------------------------------------------------------------------------
def build_conv_fwd(a_hi, a_lo, w_hi, w_lo, out, M, N, K, passes=3, geom=None, block_n=0, scale=None, bias=None,
                   relu=False, stats=None, bn=None, coef=None, counter=None, halo_mode=-1, alpha=1.0,
                   out_planes=None, ep_coef=None, res_planes=None, res_raw=None, res_coef=None, stats_only=False,
                   bn_save=False, alpha_dev=None, kchunk=0, taps=1, shift_w=0):
    """geom=None: A is [M,K]; else dict(batch,H,W,Cin,R,S,stride,pad_lo_h,pad_lo_w,pad_hi_h,pad_hi_w) (NHWC gather).
    bn/coef/counter (with stats): fuse the train-mode BatchNorm finalize of module `bn` into the kernel tail.
    stats_only: statistics pass (nothing stored; `out` may be None); 2 = the transposed form (channels along the
    accumulator rows: sums are in-register adds), available when geom is None (1x1 stride-1 convolutions).
    out_planes=(hi, lo) with ep_coef [2N]: "apply" epilogue - planes = relu?(acc*scale + shift + residual), residual =
    res_planes (hi, lo) or bn(res_raw) with res_coef; `out` must be None."""
    d = _lib.ConvDesc()
    d.a_hi = _val(a_hi, torch.float16, "a_hi")
    d.a_lo = _val(a_lo, torch.float16, "a_lo")
    d.w_hi = _val(w_hi, torch.float16, "w_hi")
    d.w_lo = _val(w_lo, torch.float16, "w_lo")
    d.out = _val(out, torch.float32, "out")
    if out is not None and kchunk == 0 and out.numel() < M * N:
        raise ValueError("conv_fwd: output buffer too small")
    if out is None and out_planes is None and not stats_only:
        raise ValueError("conv_fwd: no output")
    d.M, d.N, d.K = M, N, K
    if geom is not None:
        d.im2col = 1
        for k in ("batch", "H", "W", "Cin", "R", "S", "stride", "pad_lo_h", "pad_lo_w", "pad_hi_h", "pad_hi_w"):
            setattr(d, k, int(geom[k]))
        for k in ("a_pixel_stride", "a_row_stride", "a_img_stride"):       # 0 = dense NHWC
            setattr(d, k, int(geom.get(k, 0)))
    d.passes = passes
    d.block_n = block_n
    d.scale = _val(scale, torch.float32, "scale")
    d.bias = _val(bias, torch.float32, "bias")
    d.relu = 1 if relu else 0
    d.stats = _val(stats, torch.float64, "stats")
    d.halo_mode = halo_mode
    d.alpha = float(alpha)
    d.stats_only = int(stats_only)          # 0 | 1 statistics pass | 2 transposed statistics pass (plain GEMMs only)
    d.bn_save = 1 if bn_save else 0
    d.alpha_dev = _val(alpha_dev, torch.float32, "alpha_dev")
    d.kchunk, d.taps, d.shift_w = int(kchunk), int(taps), int(shift_w)
    if out_planes is not None:
        if out is not None or ep_coef is None:
            raise ValueError("conv_fwd: the apply epilogue writes planes only and needs ep_coef")
        hi, lo = out_planes
        if hi.numel() < M * N:
            raise ValueError("conv_fwd: output planes too small")
        d.out_hi = _val(hi, torch.float16, "out_hi")
        d.out_lo = _val(lo, torch.float16, "out_lo")
        d.ep_coef = _val(ep_coef, torch.float32, "ep_coef")
        if ep_coef.numel() < 2 * N:
            raise ValueError("conv_fwd: ep_coef must hold [2][N] floats")
        if res_planes is not None:
            d.res_kind = 1
            d.res_hi = _val(res_planes[0], torch.float16, "res_hi")
            d.res_lo = _val(res_planes[1], torch.float16, "res_lo")
        elif res_raw is not None:
            d.res_kind = 2
            d.res_raw = _val(res_raw, torch.float32, "res_raw")
            d.res_coef = _val(res_coef, torch.float32, "res_coef")
    if bn is not None:
        if stats is None or coef is None or counter is None:
            raise ValueError("conv_fwd: the fused BatchNorm finalize needs stats, coef and counter buffers")
        d.bn_gamma = _val(bn.weight, torch.float32, "bn.weight")
        d.bn_beta = _val(bn.bias, torch.float32, "bn.bias")
        d.bn_running_mean = _val(bn.running_mean, torch.float32, "bn.running_mean")
        d.bn_running_var = _val(bn.running_var, torch.float32, "bn.running_var")
        d.bn_num_batches_tracked = _val(getattr(bn, "num_batches_tracked", None), torch.int64, "bn.num_batches_tracked")
        d.bn_coef = _val(coef, torch.float32, "coef")
        d.bn_counter = _val(counter, torch.int32, "counter")
        d.bn_momentum, d.bn_eps = BN_MOMENTUM, BN_EPS
    fn = _lib.lib().vince_conv_fwd
    ref = ctypes.byref(d)
    # algorithmic FLOPs: the statistics pass of the two-pass scheme is recomputation - it gets no credit
    flops = 0.0 if stats_only else 2.0 * M * N * (geom.get("K_true", K) if geom is not None else K)
    # ... and the algorithmic HBM bytes of the launch (activations once, at fp32 width = the two fp16 planes; weights;
    # output; residual) for the launches whose role is the streaming one: statistics pass and apply epilogue
    kind = "stats" if stats_only else ("apply" if out_planes is not None else "gemm")
    in_bytes = 4.0 * (M * K if geom is None else geom["batch"] * geom["H"] * geom["W"] * geom["Cin"])
    nbytes = in_bytes + 4.0 * N * K + (0.0 if stats_only else 4.0 * M * N) + \
        (4.0 * M * N if (res_planes is not None or res_raw is not None) else 0.0)
    check = _lib.check

    def run():
        stream = ctypes.c_void_p(_stream_handle())
        if PROFILE is not None:
            # bench.py's roofline leg: CUDA events around every tensor-core launch, on the launching stream
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(fn(ref, stream), "vince_conv_fwd")
            e1.record()
            PROFILE.append((kind, flops, nbytes, e0, e1))
        else:
            check(fn(ref, stream), "vince_conv_fwd")
    run._keep = (d, a_hi, a_lo, w_hi, w_lo, out, scale, bias, stats, bn, coef, counter, out_planes, ep_coef, res_planes,
                 res_raw, res_coef, alpha_dev)
    return run


def conv_fwd(*a, **k):
    build_conv_fwd(*a, **k)()


def build_count_saturated(plane, counter):
    """counter (int64 [1], device) += number of values of the fp16 `plane` at the saturation bound +-65504 (debug aid)."""
    run = _bind(_lib.lib().vince_count_saturated, "vince_count_saturated", _ptr(plane, torch.float16, "plane"),
                plane.numel(), _ptr(counter, torch.int64, "counter"))
    run._keep = (plane, counter)
    return run


def stem_geometry(H, W):
    """Shapes of the packed stem input and the equivalent 4x1 conv (see include/vince_b200.h: vince_stem_pack)."""
    P, Q = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    Ha, Wb = P + 3, Q + 3                               # row pairs / column pairs stored (16 elements each)
    return dict(P=P, Q=Q, Ha=Ha, Wb=Wb,
                geom=dict(H=Ha, W=Q, Cin=64, R=4, S=1, stride=1, pad_lo_h=0, pad_lo_w=0, pad_hi_h=0, pad_hi_w=0,
                          a_pixel_stride=16, a_row_stride=16 * Wb, a_img_stride=16 * Wb * Ha,
                          K_true=147))                  # algorithmic K of the 7x7x3 stem (the packed K is 256)


def jigsaw_patch_size(H, W):
    """vince_model.py:145-146: both axes are padded by 3 - dim % 3 when EITHER is not a multiple of 3 (so a divisible
    axis next to a non-divisible one grows by 3); returns the patch size (PH, PW)."""
    if H % 3 != 0 or W % 3 != 0:
        H, W = H + 3 - H % 3, W + 3 - W % 3
    return H // 3, W // 3


def build_stem_pack(x, gather_idx, x_hi, x_lo, grid=1):
    """grid=3: the N frames are packed as 9N jigsaw patches (patchify folded into the stem's loads)."""
    N, C, H, W = x.shape
    if C != 3:
        raise ValueError("stem_pack: expected 3 input channels")
    run = _bind(_lib.lib().vince_stem_pack_grid, "vince_stem_pack_grid", _ptr(x, torch.float32, "x"),
                _ptr(gather_idx, torch.int64, "gather_idx"), _ptr(x_hi, torch.float16, "x_hi"),
                _ptr(x_lo, torch.float16, "x_lo"), N, H, W, grid)
    run._keep = (x, gather_idx, x_hi, x_lo)
    return run


def stem_pack(*a, **k):
    build_stem_pack(*a, **k)()


IMAGENET_MEAN = (0.485, 0.456, 0.406)      # utils/transforms.py:85,100 (== constants.py:28-29 / 255)
IMAGENET_STD = (0.229, 0.224, 0.225)


def build_stem_pack_u8(x_nhwc, gather_idx, x_hi, x_lo, mean=IMAGENET_MEAN, std=IMAGENET_STD, grid=1):
    """x_nhwc: [N,H,W,3] uint8 CUDA.  ToTensor(scale=255) + Normalize(mean, std) fused into the packing."""
    N, H, W, C = x_nhwc.shape
    if C != 3:
        raise ValueError("stem_pack_u8: expected HWC frames with 3 channels")
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    run = _bind(_lib.lib().vince_stem_pack_u8_grid, "vince_stem_pack_u8_grid", _ptr(x_nhwc, torch.uint8, "x"),
                _ptr(gather_idx, torch.int64, "gather_idx"), m3, s3, _ptr(x_hi, torch.float16, "x_hi"),
                _ptr(x_lo, torch.float16, "x_lo"), N, H, W, grid)
    run._keep = (x_nhwc, gather_idx, x_hi, x_lo, m3, s3)
    return run


def build_weight_prep(table_dev, n_entries, max_cout, w_hi, w_lo):
    """max_cout: largest Cout over the table (one block per output channel and tensor)"""
    run = _bind(_lib.lib().vince_weight_prep, "vince_weight_prep", _ptr(table_dev, torch.uint8, "table"), n_entries,
                max_cout, _ptr(w_hi, torch.float16, "w_hi"), _ptr(w_lo, torch.float16, "w_lo"))
    run._keep = (table_dev, w_hi, w_lo)
    return run


def weight_prep(*a):
    build_weight_prep(*a)()


def _residual(res_planes, res_bn):
    if res_planes is not None:
        return 1, _ptr(res_planes[0], torch.float16, "res_hi"), _ptr(res_planes[1], torch.float16, "res_lo"), None
    if res_bn is not None:
        return 2, None, None, ctypes.byref(res_bn)
    return 0, None, None, None


def build_bn_apply(main, M, C, relu, out_hi=None, out_lo=None, out_f32=None, res_planes=None, res_bn=None):
    res_kind, rh, rl, rb = _residual(res_planes, res_bn)
    run = _bind(_lib.lib().vince_bn_apply, "vince_bn_apply", ctypes.byref(main), res_kind, rh, rl, rb,
                1 if relu else 0, _ptr(out_hi, torch.float16, "out_hi"), _ptr(out_lo, torch.float16, "out_lo"),
                _ptr(out_f32, torch.float32, "out_f32"), M, C)
    run._keep = (main, res_planes, res_bn, out_hi, out_lo, out_f32)
    return run


def bn_apply(*a, **k):
    build_bn_apply(*a, **k)()


def build_bn_relu_maxpool(bn, out_hi, out_lo, N, P, Q, C):
    run = _bind(_lib.lib().vince_bn_relu_maxpool, "vince_bn_relu_maxpool", ctypes.byref(bn),
                _ptr(out_hi, torch.float16, "out_hi"), _ptr(out_lo, torch.float16, "out_lo"), N, P, Q, C)
    run._keep = (bn, out_hi, out_lo)
    return run


def bn_relu_maxpool(*a):
    build_bn_relu_maxpool(*a)()


def build_bn_final_pool(main, N, HW, C, spatial_nchw, pooled, scatter_idx=None, res_planes=None, res_bn=None):
    res_kind, rh, rl, rb = _residual(res_planes, res_bn)
    run = _bind(_lib.lib().vince_bn_final_pool, "vince_bn_final_pool", ctypes.byref(main), res_kind, rh, rl, rb,
                _ptr(scatter_idx, torch.int64, "scatter_idx"), _ptr(spatial_nchw, torch.float32, "spatial"),
                _ptr(pooled, torch.float32, "pooled"), N, HW, C)
    run._keep = (main, res_planes, res_bn, scatter_idx, spatial_nchw, pooled)
    return run


def bn_final_pool(*a, **k):
    build_bn_final_pool(*a, **k)()


def build_split_f16(x, hi, lo):
    run = _bind(_lib.lib().vince_split_f16, "vince_split_f16", _ptr(x, torch.float32, "x"),
                _ptr(hi, torch.float16, "hi"), _ptr(lo, torch.float16, "lo"), x.numel())
    run._keep = (x, hi, lo)
    return run


def split_f16(*a):
    build_split_f16(*a)()


def round_tf32(x, out):
    _lib.check(_lib.lib().vince_round_tf32(_ptr(x, torch.float32, "x"), _ptr(out, torch.float32, "out"), x.numel(),
                                          _stream()), "vince_round_tf32")


def build_l2_normalize(x, out, eps=1e-12):
    rows, D = x.shape
    run = _bind(_lib.lib().vince_l2_normalize, "vince_l2_normalize", _ptr(x, torch.float32, "x"),
                _ptr(out, torch.float32, "out"), rows, D, eps)
    run._keep = (x, out)
    return run


def l2_normalize(*a, **k):
    build_l2_normalize(*a, **k)()


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


_NCE_PLANS = {}      # (B, D, K, nf, T, device, stream) -> pre-filled descriptor, output slicing, workspace


def _nce_check(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        _ptr(t, torch.float32, name)          # raises with the precise reason
    return t.data_ptr()


def infonce_fwd(q, keys, queue_tf32, num_frames, temperature, workspace=None):
    """Fused similarity + masked multi-positive cross entropy + metrics (ONE kernel launch).  Returns a dict of device
    tensors: dists [B,nP], weights [B,nP], pos_sim [B,nP], neg_max [B], row_lse [B,2], scalars [8].
    The call is on the per-step hot path: the descriptor, the slicing of the output buffer and the workspace are cached
    per problem shape and stream, so a step costs one allocation, a dozen pointer stores and the ctypes call."""
    B, D = q.shape
    K = 0 if queue_tf32 is None else queue_tf32.shape[0]
    dev = q.device
    handle = _stream_handle()
    key = (B, D, K, num_frames, float(temperature), dev.index, handle, keys.shape[0])
    plan = _NCE_PLANS.get(key)
    if plan is None:
        if len(_NCE_PLANS) > 64:
            _NCE_PLANS.clear()
        nP = num_frames if num_frames > 0 else 1
        d = _lib.InfoNceDesc()
        d.B, d.Bk, d.K, d.D, d.num_frames = B, keys.shape[0], K, D, num_frames
        d.temperature = float(temperature)
        slices, o = [], 0
        for name, shape in (("dists", (B, nP)), ("weights", (B, nP)), ("pos_sim", (B, nP)), ("neg_max", (B,)),
                            ("row_lse", (B, 2)), ("scalars", (8,))):
            n = 1
            for s_ in shape:
                n *= s_
            slices.append((name, o, n, shape))
            o += n
        ws = torch.empty((infonce_workspace_bytes(B, D),), device=dev, dtype=torch.uint8)
        plan = _NCE_PLANS[key] = (d, ctypes.byref(d), slices, o, ws, _lib.lib().vince_infonce_fwd)
    d, ref, slices, total, ws, fn = plan
    if workspace is None:
        workspace = ws          # safe to share: launches on one stream run in order and the kernel consumes it itself
    # one allocation for all per-row outputs
    buf = torch.empty((total,), device=dev, dtype=torch.float32)
    base = buf.data_ptr()
    out = {}
    for name, o, n, shape in slices:
        out[name] = buf[o:o + n].view(shape)
        setattr(d, name, base + 4 * o)
    d.q = _nce_check(q, "q")
    d.keys = _nce_check(keys, "keys")
    d.queue_tf32 = _nce_check(queue_tf32, "queue_tf32") if K > 0 else None
    d.workspace = _val(workspace, torch.uint8, "workspace")
    _lib.check(fn(ref, ctypes.c_void_p(handle)), "vince_infonce_fwd")
    out["_workspace"] = workspace     # keep alive until the stream has consumed it
    return out


def infonce_bwd(q, keys, queue_tf32, num_frames, temperature, fwd, grad_dist=1.0, symmetric=False, dq=None,
                accumulate=False):
    """d(grad_dist * dist)/dq for the fused InfoNCE loss (what autograd produces through vince_model.py:213-233 +
    loss_util.py:7-62); `fwd` is the dict returned by infonce_fwd for the same operands.  symmetric=True is the
    self-batch loss (keys is q itself, no queue) whose columns carry gradient too.  Returns dq [B, D]."""
    B, D = q.shape
    K = 0 if queue_tf32 is None else queue_tf32.shape[0]
    dev = q.device
    if dq is None:
        if accumulate:
            raise ValueError("infonce_bwd: accumulate needs an existing dq")
        dq = torch.empty((B, D), device=dev, dtype=torch.float32)
    workspace = torch.empty((int(_lib.lib().vince_infonce_bwd_workspace_bytes(B, D)),), device=dev, dtype=torch.uint8)
    d = _lib.InfoNceDesc()
    d.q = _val(q, torch.float32, "q")
    d.keys = _val(keys, torch.float32, "keys")
    d.queue_tf32 = _val(queue_tf32, torch.float32, "queue_tf32") if K > 0 else None
    d.B, d.Bk, d.K, d.D, d.num_frames = B, keys.shape[0], K, D, num_frames
    d.temperature = float(temperature)
    d.pos_sim = _val(fwd["pos_sim"], torch.float32, "pos_sim")
    d.row_lse = _val(fwd["row_lse"], torch.float32, "row_lse")
    d.workspace = _val(workspace, torch.uint8, "workspace")
    _lib.check(_lib.lib().vince_infonce_bwd(ctypes.byref(d), float(grad_dist), 1 if symmetric else 0,
                                            1 if accumulate else 0, _ptr(dq, torch.float32, "dq"), _stream()),
               "vince_infonce_bwd")
    dq._workspace = workspace        # keep alive until the stream has consumed it
    return dq


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
