"""Host-side execution plan of the ResNet encoder forward on sm_100a kernels.

`EncoderRunner` walks a torchvision-shaped ResNet parameter tree once, records one `ConvSpec` per convolution
(where its prepared bf16 weights live, which BatchNorm follows it) and, per forward, issues

    weight_prep (1 launch, all tensors)  ->  stem_pack  ->  conv1 (im2col GEMM, +BN sums)  ->  bn+relu+maxpool
    -> per residual block:  conv (+sums) -> bn_apply(+relu) ... -> bn_apply(+residual, +relu)
    -> last block: bn_final_pool (relu(bn+residual) -> NCHW spatial features + global average pool)

Reference arithmetic being reproduced: models/building_blocks/resnet.py:76-92 (BasicBlock), :117-137 (Bottleneck),
:231-247 (stem + layers) as reached through backbone_models.py:39-54.  Activations live in HBM as NHWC bf16
(hi, lo) pairs; raw conv outputs as NHWC fp32 (see DESIGN.md "data layout").
"""
import numpy as np
import torch

from . import ops

_WEIGHT_ENTRY = np.dtype([("src", "<u8"), ("dst_off", "<i8"), ("Cout", "<i4"), ("Cin", "<i4"), ("R", "<i4"),
                          ("S", "<i4"), ("kind", "<i4"), ("pad", "<i4")])


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
    """bf16 (hi, lo) K-major copies of a set of conv / linear weights, refreshed by ONE multi-tensor launch."""

    def __init__(self, specs, passes):
        self.specs = specs
        self.passes = passes
        off = 0
        for s in specs:
            s.w_off = off
            off = _align(off + s.Cout * s.K, 64)
        self.total = off
        self.max_elems = max((s.Cout * s.K for s in specs), default=0)
        self.device = None
        self._ptrs = None
        self.w_hi = self.w_lo = self.table = None

    def _sync_device(self):
        ptrs = tuple(s.weight.data_ptr() for s in self.specs)
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
                      7 if s.kind == 1 else R, s.kind, 0)
        self.table = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        self.w_hi = torch.empty((self.total,), device=dev, dtype=torch.bfloat16)
        self.w_lo = torch.empty((self.total,), device=dev, dtype=torch.bfloat16) if self.passes == 3 else None
        self._ptrs, self.device = ptrs, dev

    def refresh(self):
        self._sync_device()
        ops.weight_prep(self.table, len(self.specs), self.max_elems, self.w_hi, self.w_lo)

    def planes(self, spec):
        n = spec.Cout * spec.K
        hi = self.w_hi[spec.w_off:spec.w_off + n]
        lo = self.w_lo[spec.w_off:spec.w_off + n] if self.w_lo is not None else None
        return hi, lo


class Act:
    """NHWC activation as bf16 planes."""
    __slots__ = ("hi", "lo", "N", "H", "W", "C")

    def __init__(self, hi, lo, N, H, W, C):
        self.hi, self.lo, self.N, self.H, self.W, self.C = hi, lo, N, H, W, C


class EncoderRunner:
    def __init__(self, model, passes=3):
        if passes not in (1, 3):
            raise ValueError("passes must be 3 (bf16x3, fp32-grade) or 1 (plain bf16)")
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
        off = 0
        for s in specs:
            s.stats_off = off
            off += 2 * s.Cout
        self.stats_total = off
        self.bank = WeightBank(specs, passes)
        self.launches = 0          # kernels launched by the last forward (bench bookkeeping)

    # ------------------------------------------------------------------------------------------
    def _conv(self, act, spec, stats):
        """raw fp32 [M, Cout] NHWC + batch sums.  Returns (raw, N, P, Q)."""
        dev = act.hi.device
        P = (act.H + 2 * spec.pad - spec.R) // spec.stride + 1
        Q = (act.W + 2 * spec.pad - spec.R) // spec.stride + 1
        M = act.N * P * Q
        raw = torch.empty((M, spec.Cout), device=dev, dtype=torch.float32)
        w_hi, w_lo = self.bank.planes(spec)
        st = stats[spec.stats_off:spec.stats_off + 2 * spec.Cout] if stats is not None else None
        if spec.R == 1 and spec.stride == 1:
            ops.conv_fwd(act.hi, act.lo, w_hi, w_lo, raw, M, spec.Cout, spec.K, passes=self.passes, stats=st)
        else:
            geom = dict(batch=act.N, H=act.H, W=act.W, Cin=act.C, R=spec.R, S=spec.R, stride=spec.stride,
                        pad_lo_h=spec.pad, pad_lo_w=spec.pad, pad_hi_h=spec.pad, pad_hi_w=spec.pad)
            ops.conv_fwd(act.hi, act.lo, w_hi, w_lo, raw, M, spec.Cout, spec.K, passes=self.passes, geom=geom, stats=st)
        self.launches += 1
        return raw, P, Q

    def _planes(self, M, C, dev):
        hi = torch.empty((M, C), device=dev, dtype=torch.bfloat16)
        lo = torch.empty((M, C), device=dev, dtype=torch.bfloat16) if self.passes == 3 else None
        return hi, lo

    def _side(self, raw, spec, stats):
        st = stats[spec.stats_off:spec.stats_off + 2 * spec.Cout] if stats is not None else None
        return ops.bn_side(raw, st, spec.bn)

    # ------------------------------------------------------------------------------------------
    def forward(self, x, train, gather_idx=None, scatter_idx=None, want_spatial=True):
        """x: [N,3,H,W] fp32 CUDA.  Returns (spatial NCHW [N,C,h,w] or None, pooled [N,C])."""
        if not x.is_cuda:
            raise RuntimeError("vince_b200 encoder: input must be a CUDA tensor (no CPU fallback)")
        if x.dtype != torch.float32:
            raise TypeError("vince_b200 encoder: input must be fp32 (the reference's arithmetic type)")
        x = x.contiguous()
        dev = x.device
        N, C3, H, W = x.shape
        self.launches = 0
        with torch.cuda.device(dev):
            self.bank.refresh()
            self.launches += 1
            stats = None
            if train:
                stats = torch.zeros((self.stats_total,), device=dev, dtype=torch.float64)
            # ---- stem ----
            sg = ops.stem_geometry(H, W)
            P, Q, Hj = sg["P"], sg["Q"], sg["Hj"]
            x_hi = torch.empty((N, Hj, Q, 64), device=dev, dtype=torch.bfloat16)
            x_lo = torch.empty_like(x_hi) if self.passes == 3 else None
            ops.stem_pack(x, gather_idx, x_hi, x_lo)
            w_hi, w_lo = self.bank.planes(self.stem)
            M = N * P * Q
            raw = torch.empty((M, 64), device=dev, dtype=torch.float32)
            st = stats[self.stem.stats_off:self.stem.stats_off + 128] if train else None
            ops.conv_fwd(x_hi, x_lo, w_hi, w_lo, raw, M, 64, 256, passes=self.passes, geom=dict(sg["geom"], batch=N),
                         stats=st)
            del x_hi, x_lo
            P2, Q2 = (P - 1) // 2 + 1, (Q - 1) // 2 + 1
            hi, lo = self._planes(N * P2 * Q2, 64, dev)
            ops.bn_relu_maxpool(self._side(raw, self.stem, stats), hi, lo, N, P, Q, 64)
            self.launches += 3
            act = Act(hi, lo, N, P2, Q2, 64)
            del raw
            # ---- residual blocks ----
            spatial = pooled = None
            for bi, blk in enumerate(self.blocks):
                last = bi == len(self.blocks) - 1
                cur = act
                convs = blk["convs"]
                for ci, spec in enumerate(convs[:-1]):
                    raw, p_, q_ = self._conv(cur, spec, stats)
                    hi, lo = self._planes(raw.shape[0], spec.Cout, dev)
                    ops.bn_apply(self._side(raw, spec, stats), raw.shape[0], spec.Cout, True, hi, lo)
                    self.launches += 1
                    cur = Act(hi, lo, cur.N, p_, q_, spec.Cout)
                spec = convs[-1]
                raw, p_, q_ = self._conv(cur, spec, stats)
                kw = {}
                if blk["down"] is not None:
                    raw_ds, _, _ = self._conv(act, blk["down"], stats)
                    kw["res_bn"] = self._side(raw_ds, blk["down"], stats)
                else:
                    kw["res_planes"] = (act.hi, act.lo)
                main = self._side(raw, spec, stats)
                if last:
                    C = spec.Cout
                    spatial = torch.empty((N, C, p_, q_), device=dev, dtype=torch.float32) if want_spatial else None
                    pooled = torch.empty((N, C), device=dev, dtype=torch.float32)
                    ops.bn_final_pool(main, N, p_ * q_, C, spatial, pooled, scatter_idx=scatter_idx, **kw)
                else:
                    hi, lo = self._planes(raw.shape[0], spec.Cout, dev)
                    ops.bn_apply(main, raw.shape[0], spec.Cout, True, hi, lo, **kw)
                    act = Act(hi, lo, N, p_, q_, spec.Cout)
                self.launches += 1
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

    def refresh(self):
        self.bank.refresh()
        self.launches = 1

    def linear(self, idx, x, relu):
        """x: [M, Cin] fp32 CUDA -> [M, Cout] fp32"""
        spec = self.specs[idx]
        M = x.shape[0]
        dev = x.device
        hi = torch.empty((M, spec.Cin), device=dev, dtype=torch.bfloat16)
        lo = torch.empty_like(hi) if self.passes == 3 else None
        ops.split_bf16(x.contiguous(), hi, lo)
        out = torch.empty((M, spec.Cout), device=dev, dtype=torch.float32)
        w_hi, w_lo = self.bank.planes(spec)
        ops.conv_fwd(hi, lo, w_hi, w_lo, out, M, spec.Cout, spec.K, passes=self.passes, bias=spec.bias, relu=relu)
        self.launches += 2
        return out
