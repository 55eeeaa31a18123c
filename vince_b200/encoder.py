"""Host-side execution plan of the ResNet encoder forward on sm_100a kernels.

`EncoderRunner` walks a torchvision-shaped ResNet parameter tree once, records one `ConvSpec` per convolution
(where its prepared fp16 weights live, which BatchNorm follows it) and, per input shape, builds a PLAN: a list of
pre-marshalled kernel launches over a statically allocated activation arena (buffers are reused as soon as their
last consumer has been issued).  Replaying a plan costs one ctypes call per kernel:

    weight_prep (1 launch, all tensors)  ->  stem_pack  ->  conv1 (im2col GEMM, +BN sums)  ->  bn+relu+maxpool
    -> per residual block:  conv (+sums) -> bn_apply(+relu) ... -> bn_apply(+residual, +relu)
    -> last block: bn_final_pool (relu(bn+residual) -> NCHW spatial features + global average pool)
  eval mode (running statistics known up front): BatchNorm scale/shift (+residual) + ReLU are folded into the producing
  convolution's epilogue, which writes the fp16 planes directly - no bn_apply pass at all (resnet.py:76-92,117-137);
  train mode can use the same epilogue as the second half of a statistics-pass + recompute-pass scheme (two_pass).

Reference arithmetic being reproduced: models/building_blocks/resnet.py:76-92 (BasicBlock), :117-137 (Bottleneck),
:231-247 (stem + layers) as reached through backbone_models.py:39-54.  Activations live in HBM as NHWC fp16
(hi, lo) pairs; raw conv outputs as NHWC fp32 (see DESIGN.md "data layout").
"""
import numpy as np
import torch

from . import ops

_WEIGHT_ENTRY = np.dtype([("src", "<u8"), ("dst_off", "<i8"), ("Cout", "<i4"), ("Cin", "<i4"), ("R", "<i4"),
                          ("S", "<i4"), ("kind", "<i4"), ("scale_log2", "<i4")])

# Weights are multiplied by 2^WEIGHT_SCALE_LOG2 (exact) before the fp16 hi/lo split so that the lo plane of O(1e-2)
# weights is a normal fp16 number; the conv kernel multiplies its accumulators by 2^-WEIGHT_SCALE_LOG2 (exact).
WEIGHT_SCALE_LOG2 = 6
WEIGHT_ALPHA = 2.0 ** -WEIGHT_SCALE_LOG2


def _align(n, a):
    return (n + a - 1) // a * a


class ConvSpec:
    __slots__ = ("weight", "bn", "Cout", "Cin", "R", "stride", "pad", "kind", "K", "w_off", "stats_off", "bias")

    def __init__(self, weight, bn, stride, pad, kind=0, bias=None):
        self.weight, self.bn, self.stride, self.pad, self.kind, self.bias = weight, bn, stride, pad, kind, bias
        if weight.dim() == 4:
            self.Cout, self.Cin, self.R = weight.shape[0], weight.shape[1], weight.shape[2]
        else:                                   # nn.Linear
            self.Cout, self.Cin, self.R = weight.shape[0], weight.shape[1], 1
        self.K = 256 if kind == 1 else self.R * self.R * self.Cin
        self.w_off = 0
        self.stats_off = 0


class WeightBank:
    """fp16 (hi, lo) K-major copies of a set of conv / linear weights, refreshed by ONE multi-tensor launch."""

    def __init__(self, specs, passes):
        self.specs = specs
        self.passes = passes
        off = 0
        for s in specs:
            s.w_off = off
            off = _align(off + s.Cout * s.K, 64)
        self.total = off
        self.max_cout = max((s.Cout for s in specs), default=0)
        self.device = None
        self._ptrs = None
        self.w_hi = self.w_lo = self.table = self._run = None
        self.generation = 0            # bumps whenever the device buffers are re-created (plans must be rebuilt)

    def key(self):
        return tuple(s.weight.data_ptr() for s in self.specs)

    def _sync_device(self):
        ptrs = self.key()
        dev = self.specs[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("vince_b200: parameters live on %s; the encoder runs on CUDA only (no CPU fallback). "
                               "Move the model with .to('cuda:N') first." % dev)
        if self._ptrs == ptrs and self.device == dev:
            return
        arr = np.zeros(len(self.specs), dtype=_WEIGHT_ENTRY)
        for i, s in enumerate(self.specs):
            if s.weight.dtype != torch.float32 or not s.weight.is_contiguous():
                raise TypeError("vince_b200: weights must be contiguous fp32")
            R = s.R
            arr[i] = (s.weight.data_ptr(), s.w_off, s.Cout, 3 if s.kind == 1 else s.Cin, 7 if s.kind == 1 else R,
                      7 if s.kind == 1 else R, s.kind, WEIGHT_SCALE_LOG2)
        self.table = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        self.w_hi = torch.empty((self.total,), device=dev, dtype=torch.float16)
        self.w_lo = torch.empty((self.total,), device=dev, dtype=torch.float16) if self.passes == 3 else None
        self._run = ops.build_weight_prep(self.table, len(self.specs), self.max_cout, self.w_hi, self.w_lo)
        self._ptrs, self.device = ptrs, dev
        self.generation += 1

    def refresh(self):
        self._sync_device()
        self._run()

    def planes(self, spec):
        n = spec.Cout * spec.K
        hi = self.w_hi[spec.w_off:spec.w_off + n]
        lo = self.w_lo[spec.w_off:spec.w_off + n] if self.w_lo is not None else None
        return hi, lo


class _Arena:
    """Static activation allocator used while a plan is being built: a freed buffer is handed to the next request it
    fits (best fit).  Safe because the plan always replays in the same order on one stream."""

    def __init__(self, device, keep=False):
        self.device = device
        self.free_list = []         # (nbytes, tensor)
        self.total = 0
        self.keep = keep            # taping mode: nothing is recycled (the backward reads every activation)

    def alloc(self, shape, dtype):
        numel = 1
        for s in shape:
            numel *= s
        nbytes = _align(max(numel, 1) * torch.empty((), dtype=dtype).element_size(), 512)
        best = None
        for i, (nb, _) in enumerate(self.free_list):
            if nb >= nbytes and (best is None or nb < self.free_list[best][0]):
                best = i
        if best is not None and self.free_list[best][0] <= 2 * nbytes + (1 << 20):
            nb, raw = self.free_list.pop(best)
        else:
            raw = torch.empty((nbytes,), device=self.device, dtype=torch.uint8)
            nb = nbytes
            self.total += nbytes
        t = raw[:numel * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)
        t._arena_raw = (nb, raw)
        return t

    def free(self, *tensors):
        if self.keep:
            return
        for t in tensors:
            if t is not None and hasattr(t, "_arena_raw"):
                self.free_list.append(t._arena_raw)


class Act:
    """NHWC activation as fp16 planes."""
    __slots__ = ("hi", "lo", "N", "H", "W", "C")

    def __init__(self, hi, lo, N, H, W, C):
        self.hi, self.lo, self.N, self.H, self.W, self.C = hi, lo, N, H, W, C


class _Plan:
    __slots__ = ("launches", "replay", "sat_counter", "stats", "idx_gather", "idx_scatter", "x_hi", "x_lo", "final", "out_shape", "arena_bytes",
                 "n_static", "tape", "shape", "last_input", "last_gather")


class EncoderRunner:
    def __init__(self, model, passes=3):
        if passes not in (1, 3):
            raise ValueError("passes must be 3 (fp16x3, fp32-grade) or 1 (plain fp16)")
        self.passes = passes
        self.model = model
        self.stem = ConvSpec(model.conv1.weight, model.bn1, 2, 3, kind=1)
        self.blocks = []
        specs = [self.stem]
        for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
            for blk in layer:
                entry = {"convs": [], "down": None}
                if hasattr(blk, "conv3"):
                    entry["convs"] = [ConvSpec(blk.conv1.weight, blk.bn1, 1, 0), ConvSpec(blk.conv2.weight, blk.bn2, blk.stride, 1),
                                      ConvSpec(blk.conv3.weight, blk.bn3, 1, 0)]
                else:
                    entry["convs"] = [ConvSpec(blk.conv1.weight, blk.bn1, blk.stride, 1), ConvSpec(blk.conv2.weight, blk.bn2, 1, 1)]
                if blk.downsample is not None:
                    entry["down"] = ConvSpec(blk.downsample[0].weight, blk.downsample[1], blk.stride, 0)
                specs += entry["convs"] + ([entry["down"]] if entry["down"] is not None else [])
                self.blocks.append(entry)
        # per conv, in one fp64 work buffer: [2*Cout] batch sums | [2*Cout] = 4*Cout fp32 (scale, shift, and - when a
        # backward will follow - batch mean, inv-std) | [1] counter
        off = 0
        for s in specs:
            s.stats_off = off
            off += 4 * s.Cout + 2        # (+1 pad: keeps every region 16-byte aligned for float4 coefficient loads)
        self.stats_total = off
        self.bank = WeightBank(specs, passes)
        self.launches = 0          # kernels launched by the last forward (bench bookkeeping)
        self.input_mean, self.input_std = ops.IMAGENET_MEAN, ops.IMAGENET_STD     # used for uint8 HWC inputs only
        self._plans = {}
        self._taping = False        # building / running a plan that keeps everything the backward needs
        self._last_unit = None      # the conv + BN unit most recently added to the plan under construction
        self.tape = None            # the plan of the last taped forward (read by vince_b200.backward.EncoderBackward)
        self.block_n_override = None
        import os
        self.halo_mode = int(os.environ.get("VINCE_B200_HALO", "-1"))   # -1 auto, 0 off, 1 force (3x3 stride-1 convs)
        # train mode: statistics pass + recompute pass for the wide 1x1 expansions (Bottleneck conv3): the raw [M,N] fp32
        # tensor and the separate BN-apply pass never touch HBM.  0 = never, 1 = where it pays in isolation or relieves HBM
        # (K <= 128: the 56x56 and 28x28 stages; B200, r02: ResNet-50 step 22.6 -> 21.6 ms), 2 = every expansion.
        self.two_pass = int(os.environ.get("VINCE_B200_TWOPASS", "1"))
        self.tstats = int(os.environ.get("VINCE_B200_TSTATS", "1"))     # 0: statistics pass in the untransposed form
        # debug aid: count the activation values that hit the fp16 saturation bound (+-65504) in every plane a forward
        # produces and warn when there are any - the (hi, lo) split clamps instead of overflowing, so a network whose
        # activations outgrow fp16's range would otherwise clip silently.  Costs one small launch per plane pair.
        self.check_saturation = int(os.environ.get("VINCE_B200_CHECK_SATURATION", "0"))
        self.saturated = 0          # values at the bound seen by the last checked forward
        # eval mode: BatchNorm(+residual)+ReLU folded into the producing convolution's epilogue (0 = separate bn_apply)
        self.fold_eval = int(os.environ.get("VINCE_B200_FOLD_EVAL", "1"))

    def __deepcopy__(self, memo):
        # plans hold raw device pointers of THIS model's parameters and arenas: a copy (VinceQueueModel deep-copies
        # the encoder, vince_model.py:576) must start from the copied parameter tree with no cached plans
        import copy
        new = EncoderRunner(copy.deepcopy(self.model, memo), self.passes)
        new.block_n_override = self.block_n_override
        new.two_pass, new.fold_eval, new.tstats = self.two_pass, self.fold_eval, self.tstats
        new.check_saturation = self.check_saturation
        new.input_mean, new.input_std = self.input_mean, self.input_std
        return new

    # ------------------------------------------------------------------------------------------
    def _block_n(self, M, N):
        if self.block_n_override is not None:
            return self.block_n_override(M, N)
        import os
        if os.environ.get("VINCE_B200_BN256") == "1" and N % 256 == 0:
            return 256
        return 0

    def _bn_args(self, spec, work, train):
        """kwargs wiring conv_fwd's fused train-mode BN finalize for `spec` into the plan's work buffer"""
        o, C = spec.stats_off, spec.Cout
        if not train:
            return {}
        return dict(stats=work[o:o + 2 * C], bn=spec.bn, coef=work[o + 2 * C:o + 4 * C].view(torch.float32),
                    counter=work[o + 4 * C:o + 4 * C + 1].view(torch.int32), bn_save=self._taping)

    def _conv_geom(self, act, spec):
        P = (act.H + 2 * spec.pad - spec.R) // spec.stride + 1
        Q = (act.W + 2 * spec.pad - spec.R) // spec.stride + 1
        geom = None
        if not (spec.R == 1 and spec.stride == 1):
            geom = dict(batch=act.N, H=act.H, W=act.W, Cin=act.C, R=spec.R, S=spec.R, stride=spec.stride,
                        pad_lo_h=spec.pad, pad_lo_w=spec.pad, pad_hi_h=spec.pad, pad_hi_w=spec.pad)
        return P, Q, act.N * P * Q, geom

    def _coef(self, spec, work):
        """[4*Cout] fp32: scale | shift | batch mean | inv-std (the last two only in taping mode)"""
        o, C = spec.stats_off, spec.Cout
        return work[o + 2 * C:o + 4 * C].view(torch.float32)

    def _build_conv(self, arena, act, spec, work, train, launches):
        """raw fp32 conv output (+ train-mode BatchNorm sums / finalize into the plan's work buffer)"""
        P, Q, M, geom = self._conv_geom(act, spec)
        raw = arena.alloc((M, spec.Cout), torch.float32)
        w_hi, w_lo = self.bank.planes(spec)
        launches.append(ops.build_conv_fwd(act.hi, act.lo, w_hi, w_lo, raw, M, spec.Cout, spec.K, passes=self.passes,
                                           geom=geom, block_n=self._block_n(M, spec.Cout), halo_mode=self.halo_mode,
                                           alpha=WEIGHT_ALPHA,
                                           **self._bn_args(spec, work, train)))
        self._last_unit = dict(spec=spec, x=act, raw=raw, coef=self._coef(spec, work), P=P, Q=Q, M=M, out=None)
        return raw, P, Q

    def _two_pass(self, spec):
        """Train mode: is conv+BN cheaper as a statistics pass + a recompute pass with the apply epilogue than as
        raw output + separate BN-apply?  True for the wide 1x1 expansions (Bottleneck conv3, N = 4K): their time is set
        by the fp32 output stream (HBM writes run at about half the copy bandwidth), not by the tensor core, so
        recomputing the small-K GEMM is cheaper than writing and re-reading the raw [M,N] tensor."""
        if self.two_pass == 0:
            return False
        kmax = 128 if self.two_pass == 1 else 512
        return spec.R == 1 and spec.stride == 1 and spec.Cout >= 4 * spec.K and spec.K <= kmax

    def _build_conv_planes(self, arena, act, spec, work, train, launches, relu, res_planes=None, res_side=None):
        """conv + BatchNorm (+ residual) (+ ReLU) -> fp16 planes, through the cheapest available route:
        eval mode       one launch, BatchNorm folded into the epilogue (coefficients from the running statistics);
        train, 2-pass   statistics pass (nothing stored) + recompute pass with the apply epilogue;
        train, default  raw fp32 output with fused statistics, then the streaming vince_bn_apply pass.
        res_side = (raw, coef): residual = bn(raw) (downsample branch).  Returns (Act planes, P, Q)."""
        P, Q, M, geom = self._conv_geom(act, spec)
        w_hi, w_lo = self.bank.planes(spec)
        C = spec.Cout
        common = dict(passes=self.passes, geom=geom, block_n=self._block_n(M, C), halo_mode=self.halo_mode,
                      alpha=WEIGHT_ALPHA)
        fused = (not train and self.fold_eval) or (train and self._two_pass(spec))
        if not fused:
            raw, P, Q = self._build_conv(arena, act, spec, work, train, launches)
            hi, lo = self._planes(arena, M, C)
            kw = {}
            if res_planes is not None:
                kw["res_planes"] = res_planes
            elif res_side is not None:
                kw["res_bn"] = ops.bn_side(*res_side)
            launches.append(ops.build_bn_apply(self._side(raw, spec, work), M, C, relu, hi, lo, **kw))
            arena.free(raw)
            out = Act(hi, lo, act.N, P, Q, C)
            self._last_unit["out"] = out
            self._sat(launches, out)
            return out, P, Q
        if train:
            # statistics pass: transposed form (channels on the accumulator rows) for plain GEMMs
            so = 2 if (geom is None and self.tstats) else 1
            launches.append(ops.build_conv_fwd(act.hi, act.lo, w_hi, w_lo, None, M, C, spec.K, stats_only=so,
                                               **common, **self._bn_args(spec, work, train)))
        hi, lo = self._planes(arena, M, C)
        kw = {}
        if res_planes is not None:
            kw["res_planes"] = res_planes
        elif res_side is not None:
            kw["res_raw"], kw["res_coef"] = res_side
        launches.append(ops.build_conv_fwd(act.hi, act.lo, w_hi, w_lo, None, M, C, spec.K, relu=relu,
                                           out_planes=(hi, lo), ep_coef=self._coef(spec, work), **common, **kw))
        out = Act(hi, lo, act.N, P, Q, C)
        # (no raw tensor on this route: such plans are never taped - forward() disables the two-pass route when taping)
        self._last_unit = dict(spec=spec, x=act, raw=None, coef=self._coef(spec, work), P=P, Q=Q, M=M, out=out)
        self._sat(launches, out)
        return out, P, Q

    def _sat(self, launches, act):
        if self.check_saturation:
            launches.append(ops.build_count_saturated(act.hi, self._sat_counter))

    def _planes(self, arena, M, C):
        hi = arena.alloc((M, C), torch.float16)
        lo = arena.alloc((M, C), torch.float16) if self.passes == 3 else None
        return hi, lo

    def _side(self, raw, spec, work):
        return ops.bn_side(raw, self._coef(spec, work))

    def _build_plan(self, N, H, W, train, dev, tape=False):
        """tape=True (train mode, a backward will follow): no activation buffer is recycled, the convolutions also
        store the batch mean / inv-std, and plan.tape records, per conv+BN unit, what the backward needs."""
        plan = _Plan()
        self._taping = bool(tape)
        arena = _Arena(dev, keep=tape)
        launches = []
        tape_blocks = []
        # BN work buffer (sums, coefficients, finalize counters); zeroed once per train-mode forward
        stats = work = torch.zeros((self.stats_total,), device=dev, dtype=torch.float64)
        plan.stats = work
        self._sat_counter = plan.sat_counter = torch.zeros((1,), device=dev, dtype=torch.int64)
        if not train:
            # eval-mode BatchNorm: coefficients from the running statistics, one tiny launch per BN layer
            for spec in self.bank.specs:
                launches.append(ops.build_bn_eval_coef(spec.bn, self._coef(spec, work)))
        plan.idx_gather = torch.zeros((N,), device=dev, dtype=torch.int64)
        plan.idx_scatter = torch.zeros((N,), device=dev, dtype=torch.int64)
        # ---- stem (stem_pack itself is bound per call: it reads the caller's tensor) ----
        sg = ops.stem_geometry(H, W)
        P, Q = sg["P"], sg["Q"]
        plan.x_hi = arena.alloc((N, sg["Ha"], sg["Wb"], 16), torch.float16)
        plan.x_lo = arena.alloc((N, sg["Ha"], sg["Wb"], 16), torch.float16) if self.passes == 3 else None
        w_hi, w_lo = self.bank.planes(self.stem)
        M = N * P * Q
        raw = arena.alloc((M, 64), torch.float32)
        launches.append(ops.build_conv_fwd(plan.x_hi, plan.x_lo, w_hi, w_lo, raw, M, 64, 256, passes=self.passes,
                                           geom=dict(sg["geom"], batch=N), alpha=WEIGHT_ALPHA,
                                           **self._bn_args(self.stem, work, train)))
        tape_stem = dict(spec=self.stem, raw=raw, coef=self._coef(self.stem, work), N=N, H=H, W=W, P=P, Q=Q)
        arena.free(plan.x_hi, plan.x_lo)
        P2, Q2 = (P - 1) // 2 + 1, (Q - 1) // 2 + 1
        hi, lo = self._planes(arena, N * P2 * Q2, 64)
        launches.append(ops.build_bn_relu_maxpool(self._side(raw, self.stem, stats), hi, lo, N, P, Q, 64))
        arena.free(raw)
        act = Act(hi, lo, N, P2, Q2, 64)
        self._sat(launches, act)
        # ---- residual blocks ----
        for bi, blk in enumerate(self.blocks):
            last = bi == len(self.blocks) - 1
            cur = act
            convs = blk["convs"]
            tb = dict(input=act, units=[], down=None, out=None, last=last)
            tape_blocks.append(tb)
            for spec in convs[:-1]:
                nxt, p_, q_ = self._build_conv_planes(arena, cur, spec, work, train, launches, relu=True)
                tb["units"].append(self._last_unit)
                if cur is not act:
                    arena.free(cur.hi, cur.lo)
                cur = nxt
            spec = convs[-1]
            down = blk["down"]
            if last:
                # last block: raw outputs feed the fused relu(bn+residual) -> NCHW + global-average-pool kernel
                raw, p_, q_ = self._build_conv(arena, cur, spec, work, train, launches)
                tb["units"].append(self._last_unit)
                if cur is not act:
                    arena.free(cur.hi, cur.lo)
                kw = {}
                raw_ds = None
                if down is not None:
                    raw_ds, _, _ = self._build_conv(arena, act, down, work, train, launches)
                    tb["down"] = self._last_unit
                    kw["res_bn"] = self._side(raw_ds, down, stats)
                else:
                    kw["res_planes"] = (act.hi, act.lo)
                plan.final = dict(main=self._side(raw, spec, stats), N=N, HW=p_ * q_, C=spec.Cout, kw=kw,
                                  keep=(raw, raw_ds, act))
                plan.out_shape = (N, spec.Cout, p_, q_)
                break
            raw_ds = None
            if down is None:
                res = dict(res_planes=(act.hi, act.lo))
            else:
                # downsample branch: raw fp32 output; its BatchNorm is applied where the residual is added (the apply
                # epilogue of the main convolution, or vince_bn_apply) - same bytes as identity planes, exact fp32 add
                raw_ds, _, _ = self._build_conv(arena, act, down, work, train, launches)
                tb["down"] = self._last_unit
                res = dict(res_side=(raw_ds, self._coef(down, work)))
            nxt, p_, q_ = self._build_conv_planes(arena, cur, spec, work, train, launches, relu=True, **res)
            tb["units"].append(self._last_unit)
            tb["out"] = nxt
            if cur is not act:
                arena.free(cur.hi, cur.lo)
            arena.free(raw_ds, act.hi, act.lo)
            act = nxt
        plan.launches = launches
        pre = stats.zero_ if train else None
        if self.check_saturation:
            counter = plan.sat_counter
            pre = (lambda: (stats.zero_(), counter.zero_())) if train else counter.zero_
        plan.replay = ops.GraphReplay(launches, pre=pre)
        plan.arena_bytes = arena.total
        plan.n_static = len(launches)
        plan.tape = dict(stem=tape_stem, blocks=tape_blocks, pool_out=tape_blocks[0]["input"]) if tape else None
        plan.shape = (N, H, W)
        self._taping = False
        return plan

    # ------------------------------------------------------------------------------------------
    def forward(self, x, train, gather_idx=None, scatter_idx=None, want_spatial=True, patch_grid=1, tape=False):
        """x: [N,3,H,W] fp32 CUDA (the reference's normalised frames), or [N,H,W,3] uint8 CUDA (raw HWC frames: the
        ToTensor(scale=255) + Normalize(self.input_mean, self.input_std) of utils/transforms.py:89-101 is then fused
        into the stem packing).  patch_grid=3 runs the trunk over the 9N jigsaw patches of the N frames
        (vince_model.py:144-155; the patchify is folded into the stem packing).
        Returns (spatial NCHW [N,C,h,w] or None, pooled [N,C]) with N -> 9N when patch_grid == 3."""
        if not x.is_cuda:
            raise RuntimeError("vince_b200 encoder: input must be a CUDA tensor (no CPU fallback)")
        raw_u8 = x.dtype == torch.uint8
        if raw_u8:
            if x.dim() != 4 or x.shape[-1] != 3:
                raise ValueError("vince_b200 encoder: uint8 input must be HWC frames [N,H,W,3]")
        elif x.dtype != torch.float32:
            raise TypeError("vince_b200 encoder: input must be fp32 NCHW (the reference's arithmetic type) or uint8 NHWC")
        x = x.contiguous()
        dev = x.device
        if raw_u8:
            N, H, W, C3 = x.shape
        else:
            N, C3, H, W = x.shape
        if patch_grid not in (1, 3):
            raise ValueError("patch_grid must be 1 or 3")
        if patch_grid == 3:
            H, W = ops.jigsaw_patch_size(H, W)
            N = 9 * N
            if gather_idx is not None or scatter_idx is not None:
                raise ValueError("the jigsaw patch path keeps frames in place (no gather / scatter index)")
        with torch.cuda.device(dev):
            self.bank.refresh()
            tape = bool(tape and train)
            two_pass = 0 if tape else self.two_pass        # the backward reads the raw tensors: no recompute route
            key = (N, H, W, bool(train), two_pass, self.tstats, self.fold_eval, tape, self.check_saturation, dev.index,
                   self.bank.generation)
            if self._plans and next(iter(self._plans))[-1] != self.bank.generation:
                self._plans.clear()                         # parameters moved: every cached pointer is stale
            plan = self._plans.get(key)
            if plan is None:
                if len(self._plans) >= 4:                   # bound the number of resident activation arenas
                    self._plans.pop(next(iter(self._plans)))
                saved_tp, self.two_pass = self.two_pass, two_pass
                try:
                    plan = self._build_plan(N, H, W, bool(train), dev, tape=tape)
                finally:
                    self.two_pass = saved_tp
                self._plans[key] = plan
            gi = si = None
            if gather_idx is not None:
                plan.idx_gather.copy_(gather_idx)
                gi = plan.idx_gather
            if scatter_idx is not None:
                plan.idx_scatter.copy_(scatter_idx)
                si = plan.idx_scatter
            if raw_u8:
                ops.build_stem_pack_u8(x, gi, plan.x_hi, plan.x_lo, self.input_mean, self.input_std, grid=patch_grid)()
            else:
                ops.build_stem_pack(x, gi, plan.x_hi, plan.x_lo, grid=patch_grid)()
            plan.replay()                                   # (zeroes the BN work buffer first in train mode)
            f = plan.final
            spatial = torch.empty(plan.out_shape, device=dev, dtype=torch.float32) if want_spatial else None
            pooled = torch.empty((N, f["C"]), device=dev, dtype=torch.float32)
            ops.build_bn_final_pool(f["main"], f["N"], f["HW"], f["C"], spatial, pooled, scatter_idx=si, **f["kw"])()
            self.launches = plan.n_static + 3 + (1 if train else 0)       # + weight_prep, stem_pack, final (+ memset)
            if self.check_saturation:
                self.saturated = int(plan.sat_counter.item())           # (debug mode: one synchronisation per forward)
                if self.saturated:
                    import warnings
                    warnings.warn("vince_b200: %d activation values reached the fp16 saturation bound (+-65504) in this "
                                  "forward and were clamped; the fp16x3 arithmetic is not fp32-grade for this network / "
                                  "input (see DESIGN.md, numerical strategy)" % self.saturated)
            if tape:
                plan.last_input, plan.last_gather = x, gi
                self.tape = plan
        return spatial, pooled


class HeadRunner:
    """Linear(+ReLU)+Linear projection heads as tcgen05 GEMMs (vince_model.py:38-49,163,171,177)."""

    def __init__(self, linears, passes=3):
        """linears: list of nn.Linear in execution order (weights prepared together)."""
        self.passes = passes
        self.linears = list(linears)
        self.specs = [ConvSpec(l.weight, None, 1, 0, bias=l.bias) for l in self.linears]
        self.bank = WeightBank(self.specs, passes)
        self.launches = 0

    def __deepcopy__(self, memo):
        import copy
        return HeadRunner(copy.deepcopy(self.linears, memo), self.passes)

    def refresh(self):
        self.bank.refresh()
        self.launches = 1

    def linear(self, idx, x, relu):
        """x: [M, Cin] fp32 CUDA -> [M, Cout] fp32"""
        spec = self.specs[idx]
        M = x.shape[0]
        dev = x.device
        hi = torch.empty((M, spec.Cin), device=dev, dtype=torch.float16)
        lo = torch.empty_like(hi) if self.passes == 3 else None
        ops.split_f16(x.contiguous(), hi, lo)
        out = torch.empty((M, spec.Cout), device=dev, dtype=torch.float32)
        w_hi, w_lo = self.bank.planes(spec)
        ops.conv_fwd(hi, lo, w_hi, w_lo, out, M, spec.Cout, spec.K, passes=self.passes, bias=spec.bias, relu=relu,
                     alpha=WEIGHT_ALPHA)
        self.launches += 2
        return out
