"""Query-encoder backward on sm_100a kernels.

What `loss.backward()` does for the reference at /root/reference/solvers/vince_solver.py:463-469 - autograd through
`F.normalize`, the projection MLP (models/vince_model.py:38-42,177-180), the average pool, and every conv + train-mode
BatchNorm (+ residual) (+ ReLU) unit of the ResNet trunk (models/building_blocks/resnet.py:76-92,117-137,231-247) -
replayed from the tape the taped forward left behind (encoder.py, `tape=True`):

  per unit, last to first:
    vince_bn_bwd          ReLU mask, BatchNorm backward (d gamma, d beta, dRaw), dRaw as power-of-two-scaled fp16 planes
                          (zero-dilated onto the input grid for stride-2 convolutions), masked dZ for the skip path
    weight gradient       vince_transpose_pad (dRaw and the unit's input activation -> [C][padded pixels]) +
                          vince_conv_fwd as a batched split-K GEMM (one batch per filter tap) + vince_wgrad_reduce
    data gradient         vince_conv_fwd: stride-1 convolution of dRaw with the flipped / transposed filter
                          (vince_weight_prep kind 2), fp32 output
  stem: vince_maxpool_bwd -> vince_bn_bwd (fp32 dRaw) -> vince_stem_wgrad.

Gradients land in `param.grad` (fp32, accumulated if already present) of the query encoder's parameters, so
`torch.optim.SGD(model.parameters(), ...)` of the reference solver - or `vince_b200.optim.FusedSGD` - steps unchanged.
No autograd graph, no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from .encoder import WEIGHT_ALPHA, WEIGHT_SCALE_LOG2, Act, WeightBank, _align, _Arena

_lib_check = _lib.check


class _DgradSpec:
    """weight_prep kind 2 entry: rows = Cin, K = R*S*Cout (flipped, channel-transposed filter)."""
    __slots__ = ("weight", "Cout", "Cin", "R", "kind", "K", "w_off", "src")

    def __init__(self, src):
        self.src = src
        self.weight = src.weight
        # WeightBank sizes the planes with Cout * K and launches one block per "Cout" row: here rows = Cin
        self.Cout, self.Cin, self.R, self.kind = src.Cin, src.Cout, src.R, 2
        self.K = src.R * src.R * src.Cout
        self.w_off = 0


class _DgradBank(WeightBank):
    def _sync_device(self):
        # the table of kind-2 entries carries the ORIGINAL (Cout, Cin); rows (= our .Cout) is only the launch extent
        ptrs = self.key()
        dev = self.specs[0].weight.device
        if self._ptrs == ptrs and self.device == dev:
            return
        from .encoder import _WEIGHT_ENTRY
        arr = np.zeros(len(self.specs), dtype=_WEIGHT_ENTRY)
        for i, s in enumerate(self.specs):
            arr[i] = (s.weight.data_ptr(), s.w_off, s.src.Cout, s.src.Cin, s.R, s.R, 2, WEIGHT_SCALE_LOG2)
        self.table = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        self.w_hi = torch.empty((self.total,), device=dev, dtype=torch.float16)
        self.w_lo = torch.empty((self.total,), device=dev, dtype=torch.float16) if self.passes == 3 else None
        self._run = ops.build_weight_prep(self.table, len(self.specs), self.max_cout, self.w_hi, self.w_lo)
        self._ptrs, self.device = ptrs, dev
        self.generation += 1


def _p(t, dtype=None):
    return ops._ptr(t, dtype, "tensor") if t is not None else None


def build_bn_bwd(dA, dB, raw, coef, M, C, work, mask_kind=0, out_planes=None, bcast_hw=0, dgamma=None, dbeta=None,
                 accumulate=False, d_planes=None, d_f32=None, dz_out=None, dil=1, geom=None):
    """validated, pre-marshalled vince_bn_bwd launch (zero-argument callable)"""
    d = _lib.BnBwdDesc()
    d.dA, d.dB = ops._val(dA, torch.float32, "dA"), ops._val(dB, torch.float32, "dB")
    d.bcast_hw, d.mask_kind = int(bcast_hw), int(mask_kind)
    if out_planes is not None:
        d.out_hi = ops._val(out_planes[0], torch.float16, "out_hi")
        d.out_lo = ops._val(out_planes[1], torch.float16, "out_lo")
    d.raw, d.coef = ops._val(raw, torch.float32, "raw"), ops._val(coef, torch.float32, "coef")
    d.M, d.C = int(M), int(C)
    d.work = ops._val(work, torch.float64, "work")
    d.dgamma, d.dbeta = ops._val(dgamma, torch.float32, "dgamma"), ops._val(dbeta, torch.float32, "dbeta")
    d.accumulate = 1 if accumulate else 0
    if d_planes is not None:
        d.d_hi = ops._val(d_planes[0], torch.float16, "d_hi")
        d.d_lo = ops._val(d_planes[1], torch.float16, "d_lo")
    d.d_f32, d.dz_out = ops._val(d_f32, torch.float32, "d_f32"), ops._val(dz_out, torch.float32, "dz_out")
    d.dil = int(dil)
    if geom is not None:
        d.P, d.Q, d.Hd, d.Wd = [int(v) for v in geom]
    run = ops._bind(_lib.lib().vince_bn_bwd, "vince_bn_bwd", ctypes.byref(d))
    run._keep = (d, dA, dB, raw, coef, work, out_planes, dgamma, dbeta, d_planes, d_f32, dz_out)
    return run


def bn_bwd(*a, **k):
    build_bn_bwd(*a, **k)()


def build_transpose_pad(src, dst, M, C, P, Q, stride, offset, Hp, Wp, ld, copies=1):
    run = ops._bind(_lib.lib().vince_transpose_pad, "vince_transpose_pad", _p(src[0], torch.float16),
                    _p(src[1], torch.float16), _p(dst[0], torch.float16), _p(dst[1], torch.float16), M, C, P, Q, stride,
                    offset, Hp, Wp, ld, copies)
    run._keep = (src, dst)
    return run


def transpose_pad(*a, **k):
    build_transpose_pad(*a, **k)()


def sgemm(A, B, C, M, N, K, lda, ldb, ldc, ta=False, tb=False, accumulate=False, relu_mask=None):
    _lib_check(_lib.lib().vince_sgemm(_p(A, torch.float32), _p(B, torch.float32), _p(C, torch.float32), M, N, K, lda, ldb,
                                      ldc, 1 if ta else 0, 1 if tb else 0, 1 if accumulate else 0,
                                      _p(relu_mask, torch.float32), ops._stream()), "vince_sgemm")


def colsum(x, out, accumulate=False):
    R, C = x.shape
    _lib_check(_lib.lib().vince_colsum(_p(x, torch.float32), _p(out, torch.float32), R, C, 1 if accumulate else 0,
                                       ops._stream()), "vince_colsum")


class GradSlots:
    """Assigns `.grad` tensors as views of ONE flat fp32 buffer (a single NCCL all-reduce covers every gradient) and
    tracks which have been written during the current backward (first write overwrites, later writes accumulate)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += _align(p.numel(), 4)
        self.total = off
        self.flat = None
        self.written = set()

    def begin(self, device):
        if self.flat is None or self.flat.device != device:
            self.flat = torch.zeros((self.total,), device=device, dtype=torch.float32)
        self.written = set()
        self.fresh = set()
        for p in self.params:
            if p.grad is None:
                o = self.offsets[id(p)]
                p.grad = self.flat[o:o + p.numel()].view(p.shape)
                self.fresh.add(id(p))

    def target(self, p):
        """(grad tensor, accumulate?) for a parameter about to receive its gradient"""
        acc = id(p) not in self.fresh or id(p) in self.written
        self.written.add(id(p))
        return p.grad, acc

    def end(self):
        # parameters that took no part in the loss (torchvision's unused `fc`): autograd leaves their .grad None
        for p in self.params:
            if id(p) in self.fresh and id(p) not in self.written:
                p.grad = None


class EncoderBackward:
    """Backward of one EncoderRunner (ResNet trunk) from its tape."""

    def __init__(self, runner):
        self.runner = runner
        specs = [_DgradSpec(s) for s in runner.bank.specs if s is not runner.stem]
        off = 0
        for s in specs:
            s.w_off = off
            off = _align(off + s.Cout * s.K, 64)
        self.dbank = _DgradBank.__new__(_DgradBank)
        WeightBank.__init__(self.dbank, specs, runner.passes)
        self.dspec = {id(s.src): s for s in specs}
        # weight gradients are leaves: a rounding error there does not propagate, and the 2^-12 relative error of a
        # single fp16 pass averages out over the 10^4..10^6 pixel products of one gradient element - one pass (hi planes
        # only: a third of the MMAs, half of the transposed bytes) unless VINCE_B200_WGRAD_PASSES=3
        import os
        self.wgrad_passes = int(os.environ.get("VINCE_B200_WGRAD_PASSES", "1"))
        self.launches = 0
        # a backward is a static launch list for a given forward plan: built (and run) once, then replayed - one ctypes
        # call per kernel instead of ~40 us of Python per op (the first version was host-bound: 38 ms per ResNet-18
        # step for 31 ms of kernels)
        self._replay = {}
        self._rec = None

    def _do(self, fn):
        fn()
        if self._rec is not None:
            self._rec.append(fn)

    # ---- one conv unit: weight gradient + (optionally) data gradient from dRaw planes on the unit's INPUT grid ----
    def _conv_grads(self, arena, u, dplanes, scale2, slots, need_dx):
        spec, x = u["spec"], u["x"]
        N, H, W, Cin, Cout, R, pad = x.N, x.H, x.W, x.C, spec.Cout, spec.R, spec.pad
        dev = x.hi.device
        Min = N * H * W
        # -- weight gradient: batched split-K GEMM over the (padded) pixel axis --
        # padded row pitch rounded to 8 elements: a tap's row shift is then a 16-byte aligned TMA coordinate; its
        # column shift (-1 / 0 / +1) picks one of three pre-shifted copies of the transposed activation
        Hp = H + 2 * pad
        Wp = _align(W + 2 * pad, 8) if R > 1 else W
        ld = _align(N * Hp * Wp, 8)
        taps = R * R
        copies = 3 if taps == 9 else 1
        two = self.wgrad_passes == 3
        xt = (arena.alloc((copies * Cin, ld), torch.float16), arena.alloc((copies * Cin, ld), torch.float16) if two else None)
        dt = (arena.alloc((Cout, ld), torch.float16), arena.alloc((Cout, ld), torch.float16) if two else None)
        if pad > 0 or ld != N * Hp * Wp:               # holes (padding ring / row-pitch slack) must read as zero
            for t in xt + dt:
                if t is not None:
                    self._do(t.zero_)
        self._do(build_transpose_pad((x.hi, x.lo if two else None), xt, Min, Cin, H, W, 1, pad, Hp, Wp, ld, copies=copies))
        self._do(build_transpose_pad((dplanes[0], dplanes[1] if two else None), dt, Min, Cout, H, W, 1, pad, Hp, Wp, ld))
        mpad = _align(Cout, 128)
        kblocks = (ld + 63) // 64
        base_tiles = taps * (mpad // 128) * max(1, Cin // 128)
        splits = max(1, min((148 * 6 + base_tiles - 1) // base_tiles, (kblocks + 15) // 16))
        kchunk = ((kblocks + splits - 1) // splits) * 64
        splits = (ld + kchunk - 1) // kchunk
        part = arena.alloc((taps * splits * mpad, Cin), torch.float32)
        self._do(ops.build_conv_fwd(dt[0], dt[1], xt[0], xt[1], part, Cout, Cin, ld, passes=self.wgrad_passes,
                                    kchunk=kchunk, taps=taps, shift_w=Wp))
        grad, acc = slots.target(spec.weight)
        inv = scale2[1:2]
        red = ops._bind(_lib.lib().vince_wgrad_reduce, "vince_wgrad_reduce", _p(part, torch.float32), taps, splits, mpad,
                        Cout, Cin, _p(inv, torch.float32), 1.0, _p(grad, torch.float32), 1 if acc else 0)
        red._keep = (part, inv, grad)
        self._do(red)
        arena.free(part, *[t for t in xt + dt if t is not None])
        self.launches += 8
        if not need_dx:
            return None
        # -- data gradient: stride-1 convolution with the flipped / transposed filter --
        ds = self.dspec[id(spec)]
        w_hi, w_lo = self.dbank.planes(ds)
        dx = arena.alloc((Min, Cin), torch.float32)
        geom = None
        if R > 1:
            geom = dict(batch=N, H=H, W=W, Cin=Cout, R=R, S=R, stride=1, pad_lo_h=R - 1 - pad, pad_lo_w=R - 1 - pad,
                        pad_hi_h=R - 1 - pad, pad_hi_w=R - 1 - pad)
        self._do(ops.build_conv_fwd(dplanes[0], dplanes[1], w_hi, w_lo, dx, Min, Cin, ds.K, passes=self.runner.passes,
                                    geom=geom, alpha=WEIGHT_ALPHA, alpha_dev=scale2[1:2],
                                    halo_mode=self.runner.halo_mode))
        self.launches += 1
        return dx

    def _unit(self, arena, u, dA, dB, slots, mask_kind, out_planes=None, bcast_hw=0, want_dz=False, need_dx=True):
        """backward of one conv + BN (+ReLU) unit; returns (dX fp32 [N*H*W, Cin] or None, dZ or None)"""
        spec, x = u["spec"], u["x"]
        C, M = spec.Cout, u["M"]
        dev = x.hi.device
        work = arena.alloc((3 * C + 2,), torch.float64)
        st = spec.stride
        if st == 1:
            dplanes = (arena.alloc((M, C), torch.float16), arena.alloc((M, C), torch.float16))
            geom = None
        else:
            # gradient planes live on the INPUT grid (zero-dilated): both the data gradient (a stride-1 convolution)
            # and the weight gradient (tap shifts along the padded pixel axis) then see a stride-1 geometry
            dplanes = (arena.alloc((x.N * x.H * x.W, C), torch.float16), arena.alloc((x.N * x.H * x.W, C), torch.float16))
            self._do(dplanes[0].zero_)
            self._do(dplanes[1].zero_)
            geom = (u["P"], u["Q"], x.H, x.W)
        dz = arena.alloc((M, C), torch.float32) if want_dz else None
        gw, accw = slots.target(spec.bn.weight)
        gb, accb = slots.target(spec.bn.bias)
        self._do(build_bn_bwd(dA, dB, u["raw"], u["coef"], M, C, work, mask_kind=mask_kind, out_planes=out_planes,
                              bcast_hw=bcast_hw, dgamma=gw, dbeta=gb, accumulate=accw, d_planes=dplanes, dz_out=dz,
                              dil=st, geom=geom))
        self.launches += 4
        scale2 = work[3 * C:3 * C + 1].view(torch.float32)                 # [2^e, 2^-e]
        dx = self._conv_grads(arena, u, dplanes, scale2, slots, need_dx)
        arena.free(work, *dplanes)
        return dx, dz

    def _signature(self, plan, slots, d_pooled):
        sig = [id(plan), d_pooled.shape[0], d_pooled.shape[1]]
        for s_ in self.runner.bank.specs:
            for p in (s_.weight, s_.bn.weight, s_.bn.bias):
                sig.append(p.grad.data_ptr() if p.grad is not None else 0)
                sig.append(id(p) in slots.fresh and id(p) not in slots.written)      # first write of this step?
        return tuple(sig)

    def run(self, d_pooled, slots):
        """d_pooled: [N, C] fp32 gradient wrt the pooled features in the encoder's INTERNAL (shuffled) row order.
        Writes the gradients of every trunk parameter (conv weights, BatchNorm gamma / beta)."""
        plan = self.runner.tape
        if plan is None or plan.tape is None:
            raise RuntimeError("vince_b200 backward: no taped forward to differentiate (run the model in train mode "
                               "with gradients enabled first)")
        dev = d_pooled.device
        with torch.cuda.device(dev):
            self.dbank.refresh()
            sig = self._signature(plan, slots, d_pooled)
            cached = self._replay.get(id(plan))
            if cached is not None and cached["sig"] == sig:
                cached["d_in"].copy_(d_pooled)
                cached["graph"]()                           # the recorded launch list, as one CUDA graph launch
                self._stem_wgrad(plan, cached["draw"], cached["stem_grad"], cached["stem_acc"])
                for p in cached["params"]:
                    slots.written.add(id(p))
                self.launches = cached["n"]
                return
            self._rec = []
            written_before = set(slots.written)
            d_in = torch.empty_like(d_pooled)
            d_in.copy_(d_pooled)
            try:
                draw, stem_grad, stem_acc = self._build_and_run(plan, d_in, slots)
            finally:
                rec, self._rec = self._rec, None
            params = [p for p in slots.params if id(p) in slots.written and id(p) not in written_before]
            self._replay = {id(plan): dict(sig=sig, launches=rec, graph=ops.GraphReplay(rec), d_in=d_in, draw=draw, stem_grad=stem_grad,
                                           stem_acc=stem_acc, params=params, n=self.launches, plan=plan)}

    def _stem_wgrad(self, plan, draw, grad, acc):
        # bound per call: reads the caller's input tensor of the last forward (prefetch buffers alternate)
        stem = plan.tape["stem"]
        x = plan.last_input
        is_u8 = x.dtype == torch.uint8
        m3 = (ctypes.c_float * 3)(*[float(v) for v in self.runner.input_mean])
        s3 = (ctypes.c_float * 3)(*[float(v) for v in self.runner.input_std])
        _lib_check(_lib.lib().vince_stem_wgrad(None if is_u8 else _p(x, torch.float32), _p(x, torch.uint8) if is_u8 else None,
                                               _p(plan.last_gather, torch.int64), m3, s3, _p(draw, torch.float32),
                                               _p(grad, torch.float32), stem["N"], stem["H"], stem["W"], 1 if acc else 0,
                                               ops._stream()), "vince_stem_wgrad")

    def _build_and_run(self, plan, d_pooled, slots):
        tape = plan.tape
        dev = d_pooled.device
        self.launches = 0
        if True:
            arena = _Arena(dev)
            dA, dB = d_pooled, None
            for bi in range(len(tape["blocks"]) - 1, -1, -1):
                blk = tape["blocks"][bi]
                units, down = blk["units"], blk["down"]
                last_u = units[-1]
                if blk["last"]:
                    # relu(bn(main) + residual) -> average pool: the mask is the sign of the block output, recomputed
                    # from the raw tensors (the final kernel never wrote planes)
                    out_planes = self._final_block_planes(arena, plan)
                    bhw = last_u["P"] * last_u["Q"]
                else:
                    out_planes, bhw = (blk["out"].hi, blk["out"].lo), 0
                dx, dz = self._unit(arena, last_u, dA, dB, slots, mask_kind=2, out_planes=out_planes, bcast_hw=bhw,
                                    want_dz=True)
                if blk["last"]:
                    arena.free(*out_planes)
                arena.free(dA, dB)                  # (d_pooled is not an arena buffer: ignored)
                for u in reversed(units[:-1]):
                    ndx, _ = self._unit(arena, u, dx, None, slots, mask_kind=1)
                    arena.free(dx)
                    dx = ndx
                if down is not None:
                    ddx, _ = self._unit(arena, down, dz, None, slots, mask_kind=0)
                    arena.free(dz)
                    dA, dB = dx, ddx
                else:
                    dA, dB = dx, dz
            # ---- max pool + stem ----
            stem = tape["stem"]
            N, P, Q = stem["N"], stem["P"], stem["Q"]
            dpool = arena.alloc((N * P * Q, 64), torch.float32)
            mp = ops._bind(_lib.lib().vince_maxpool_bwd, "vince_maxpool_bwd", _p(dA, torch.float32), _p(dB, torch.float32),
                           _p(stem["raw"], torch.float32), _p(stem["coef"], torch.float32), _p(dpool, torch.float32),
                           N, P, Q, 64)
            mp._keep = (dA, dB, dpool)
            self._do(mp)
            arena.free(dA, dB)
            work = arena.alloc((3 * 64 + 2,), torch.float64)
            draw = arena.alloc((N * P * Q, 64), torch.float32)
            spec = stem["spec"]
            gw, accw = slots.target(spec.bn.weight)
            gb, accb = slots.target(spec.bn.bias)
            self._do(build_bn_bwd(dpool, None, stem["raw"], stem["coef"], N * P * Q, 64, work, mask_kind=1, dgamma=gw,
                                  dbeta=gb, accumulate=accw, d_f32=draw))
            grad, acc = slots.target(spec.weight)
            self._stem_wgrad(plan, draw, grad, acc)
            self.launches += 8
            return draw, grad, acc

    def _final_block_planes(self, arena, plan):
        """relu(bn(main) + residual) of the last block as planes (only its sign is used): the forward's final kernel
        wrote NCHW fp32 + the pool, so the NHWC planes are recomputed with one vince_bn_apply."""
        f = plan.final
        M, C = f["N"] * f["HW"], f["C"]
        hi, lo = arena.alloc((M, C), torch.float16), arena.alloc((M, C), torch.float16)
        self._do(ops.build_bn_apply(f["main"], M, C, True, hi, lo, **f["kw"]))
        self.launches += 1
        return hi, lo


class HeadBackward:
    """F.normalize + Linear/ReLU/Linear backward (vince_model.py:38-42,177-180) with plain fp32 GEMMs."""

    @staticmethod
    def run(linears, saved, dq, slots, eps=1e-12):
        """saved: dict(pooled [B,C], hidden [B,C], prenorm [B,D]); dq: [B,D] gradient wrt the normalised embeddings.
        Returns d_pooled [B,C]."""
        l1, l2 = linears
        pooled, hidden, prenorm = saved["pooled"], saved["hidden"], saved["prenorm"]
        B, C = pooled.shape
        D = prenorm.shape[1]
        dev = dq.device
        with torch.cuda.device(dev):
            dpre = torch.empty((B, D), device=dev, dtype=torch.float32)
            _lib_check(_lib.lib().vince_normalize_bwd(_p(prenorm, torch.float32), _p(dq.contiguous(), torch.float32),
                                                      _p(dpre, torch.float32), B, D, eps, 1.0, ops._stream()),
                       "vince_normalize_bwd")
            # Linear 2: prenorm = hidden @ W2^T + b2,  W2 [D, C]
            g, acc = slots.target(l2.weight)
            sgemm(dpre, hidden, g, D, C, B, D, C, C, ta=True, accumulate=acc)                 # dW2 = dpre^T hidden
            g, acc = slots.target(l2.bias)
            colsum(dpre, g, acc)
            dhid = torch.empty((B, C), device=dev, dtype=torch.float32)
            sgemm(dpre, l2.weight, dhid, B, C, D, D, C, C, relu_mask=hidden)                   # dhid = (dpre W2) * [hidden > 0]
            # Linear 1: hidden = relu(pooled @ W1^T + b1),  W1 [C, C]
            g, acc = slots.target(l1.weight)
            sgemm(dhid, pooled, g, C, C, B, C, C, C, ta=True, accumulate=acc)
            g, acc = slots.target(l1.bias)
            colsum(dhid, g, acc)
            dpool = torch.empty((B, C), device=dev, dtype=torch.float32)
            sgemm(dhid, l1.weight, dpool, B, C, C, C, C, C)
        return dpool
