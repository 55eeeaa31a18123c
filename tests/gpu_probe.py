"""Diagnostic script for a GPU box: exercises every kernel once against torch fp64 references and prints one line
per check (continues after failures, so one gpurun call yields a full picture).  Not collected by pytest.

    python tests/gpu_probe.py [name-filter ...]
"""
import os
import sys
import traceback

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from vince_b200 import ops  # noqa: E402
import vince_oracle as vo  # noqa: E402

DEV = "cuda"
RESULTS = []


def split(x):
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    return hi.contiguous(), lo.contiguous()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def report(name, err, tol, extra=""):
    ok = err <= tol and np.isfinite(err)
    RESULTS.append((name, ok))
    print("%-44s %s err=%.3e tol=%.1e %s" % (name, "PASS" if ok else "FAIL", err, tol, extra), flush=True)


def weight_table(entries):
    """entries: list of (src_tensor, dst_off, Cout, Cin, R, S, kind) -> uint8 device tensor."""
    dt = np.dtype([("src", "<u8"), ("dst_off", "<i8"), ("Cout", "<i4"), ("Cin", "<i4"), ("R", "<i4"), ("S", "<i4"),
                   ("kind", "<i4"), ("scale_log2", "<i4")])
    arr = np.zeros(len(entries), dtype=dt)
    for i, (src, off, co, ci, r, s, kind) in enumerate(entries):
        arr[i] = (src.data_ptr(), off, co, ci, r, s, kind, 0)
    return torch.from_numpy(arr.view(np.uint8).copy()).to(DEV)


def prep_weight(w, kind=0):
    Cout, Cin, R, S = w.shape
    K = 256 if kind == 1 else R * S * Cin
    hi = torch.empty((Cout, K), device=DEV, dtype=torch.float16)
    lo = torch.empty_like(hi)
    tab = weight_table([(w, 0, Cout, Cin, R, S, kind)])
    ops.weight_prep(tab, 1, Cout, hi, lo)
    torch.cuda.synchronize()
    return hi, lo


# ----------------------------------------------------------------------------------------------------------
def check_gemm_tiled():
    g = torch.Generator().manual_seed(0)
    for (M, N, K, passes, bn) in [(300, 128, 256, 3, 0), (300, 128, 256, 1, 0), (8, 64, 512, 3, 0), (1000, 256, 128, 3, 256),
                                  (256, 512, 512, 3, 128), (130, 32, 64, 3, 0)]:
        a = torch.randn((M, K), generator=g).to(DEV)
        w = (torch.randn((N, K), generator=g) * 0.05).to(DEV)
        a_hi, a_lo = split(a)
        w_hi, w_lo = split(w)
        out = torch.full((M, N), float("nan"), device=DEV)
        stats = torch.zeros((2, N), device=DEV, dtype=torch.float64)
        bnm = _BN(N, g)
        rm0, rv0 = bnm.running_mean.clone(), bnm.running_var.clone()
        coef = torch.zeros((2 * N,), device=DEV)
        counter = torch.zeros((2,), device=DEV, dtype=torch.int32)
        ops.conv_fwd(a_hi, a_lo if passes == 3 else None, w_hi, w_lo if passes == 3 else None, out, M, N, K,
                     passes=passes, block_n=bn, stats=stats.view(-1), bn=bnm, coef=coef, counter=counter)
        torch.cuda.synchronize()
        ref = a.double() @ w.double().t()
        tol = 3e-5 if passes == 3 else 1e-2
        report("gemm_tiled M%d N%d K%d p%d bn%d" % (M, N, K, passes, bn), rel(out, ref), tol)
        report("  stats sum", rel(stats[0], ref.sum(0)), 1e-4 if passes == 3 else 2e-2)
        report("  stats sumsq", rel(stats[1], (ref * ref).sum(0)), 1e-4 if passes == 3 else 2e-2)
        if passes == 3:
            sc_ref, sh_ref, rm_new, rv_new = bn_coef_ref(out, bnm, True, rm0, rv0)
            report("  fused bn coef", rel(coef[:N], sc_ref) + rel(coef[N:], sh_ref), 1e-5)
            report("  fused bn running", rel(bnm.running_mean, rm_new) + rel(bnm.running_var, rv_new), 1e-5,
                   "nbt=%d" % int(bnm.num_batches_tracked))
    # bias + relu + scale epilogue
    M, N, K = 200, 128, 128
    a = torch.randn((M, K), generator=g).to(DEV)
    w = (torch.randn((N, K), generator=g) * 0.1).to(DEV)
    bias = torch.randn((N,), generator=g).to(DEV)
    scale = (torch.rand((N,), generator=g) + 0.5).to(DEV)
    a_hi, a_lo = split(a)
    w_hi, w_lo = split(w)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.conv_fwd(a_hi, a_lo, w_hi, w_lo, out, M, N, K, passes=3, scale=scale, bias=bias, relu=True)
    torch.cuda.synchronize()
    ref = F.relu((a.double() @ w.double().t()) * scale.double() + bias.double())
    report("gemm_tiled scale+bias+relu", rel(out, ref), 3e-5)


def conv_case(name, batch, H, W, Cin, Cout, R, stride, pad, passes=3, bn=0, halo=0):
    g = torch.Generator().manual_seed(abs(hash(name)) % 1000)
    x = torch.randn((batch, Cin, H, W), generator=g).to(DEV)
    w = (torch.randn((Cout, Cin, R, R), generator=g) * (2.0 / (Cin * R * R)) ** 0.5).to(DEV)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    a_hi, a_lo = split(x_nhwc)
    w_hi, w_lo = prep_weight(w)
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    M = batch * P * Q
    out = torch.full((M, Cout), float("nan"), device=DEV)
    stats = torch.zeros((2, Cout), device=DEV, dtype=torch.float64)
    geom = dict(batch=batch, H=H, W=W, Cin=Cin, R=R, S=R, stride=stride, pad_lo_h=pad, pad_lo_w=pad, pad_hi_h=pad,
                pad_hi_w=pad)
    ops.conv_fwd(a_hi, a_lo if passes == 3 else None, w_hi, w_lo if passes == 3 else None, out, M, Cout, R * R * Cin,
                 passes=passes, geom=geom, block_n=bn, stats=stats.view(-1), halo_mode=halo)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), stride=stride, padding=pad).permute(0, 2, 3, 1).reshape(M, Cout)
    e = rel(out, ref)
    report(name, e, 3e-5 if passes == 3 else 1e-2)
    if e > 1e-3 and passes == 3:
        # localise: per-row error pattern
        d = (out.double().cpu() - ref.cpu()).abs().amax(1)
        bad = (d > 1e-3 * ref.abs().max().item()).nonzero().flatten()
        print("    bad rows: %d of %d; first %s" % (bad.numel(), M, bad[:16].tolist()))
    report("  stats", rel(stats[0], ref.sum(0)) + rel(stats[1], (ref * ref).sum(0)), 2e-4 if passes == 3 else 5e-2)


def check_conv_im2col():
    conv_case("conv3x3 s1 14x14 64->64", 2, 14, 14, 64, 64, 3, 1, 1)
    conv_case("conv3x3 s1 56x56 64->64 b3", 3, 56, 56, 64, 64, 3, 1, 1)
    conv_case("conv3x3 s2 28x28 64->128", 2, 28, 28, 64, 128, 3, 2, 1)
    conv_case("conv1x1 s2 28x28 64->128", 2, 28, 28, 64, 128, 1, 2, 0)
    conv_case("conv3x3 s1 7x7 512->512", 4, 7, 7, 512, 512, 3, 1, 1)
    conv_case("conv3x3 s2 15x15 128->256 (odd)", 3, 15, 15, 128, 256, 3, 2, 1)
    conv_case("conv1x1 s1 as im2col 9x9 128->64", 2, 9, 9, 128, 64, 1, 1, 0)
    conv_case("conv3x3 s1 fp16x1", 2, 14, 14, 64, 64, 3, 1, 1, passes=1)
    conv_case("conv3x3 s1 14x14 256->256 bn256", 4, 14, 14, 256, 256, 3, 1, 1, bn=256)


def check_conv_halo():
    conv_case("halo 56x56 64->64 b3", 3, 56, 56, 64, 64, 3, 1, 1, halo=1)
    conv_case("halo 28x28 128->128 b5", 5, 28, 28, 128, 128, 3, 1, 1, halo=1)
    conv_case("halo 14x14 256->256 b4", 4, 14, 14, 256, 256, 3, 1, 1, halo=1)
    conv_case("halo 7x7 512->512 b4", 4, 7, 7, 512, 512, 3, 1, 1, halo=1)
    conv_case("halo 15x13 64->128 (odd)", 3, 15, 13, 64, 128, 3, 1, 1, halo=1)
    conv_case("halo 19x19 128->64 b2", 2, 19, 19, 128, 64, 3, 1, 1, halo=1)
    conv_case("halo 8x126 64->64 b2 (Wp=128)", 2, 8, 126, 64, 64, 3, 1, 1, halo=1)
    conv_case("halo 56x56 fp16x1", 2, 56, 56, 64, 64, 3, 1, 1, passes=1, halo=1)
    conv_case("halo auto 56x56 64->64 b16", 16, 56, 56, 64, 64, 3, 1, 1, halo=-1)


def check_stem():
    g = torch.Generator().manual_seed(5)
    for (N, H, W) in [(2, 64, 64), (3, 75, 51), (2, 224, 224), (2, 225, 225)]:
        x = torch.randn((N, 3, H, W), generator=g).to(DEV)
        w = (torch.randn((64, 3, 7, 7), generator=g) * 0.1).to(DEV)
        perm = torch.randperm(N, generator=g).to(DEV)
        sg = ops.stem_geometry(H, W)
        P, Q = sg["P"], sg["Q"]
        x_hi = torch.empty((N, sg["Ha"], sg["Wb"], 16), device=DEV, dtype=torch.float16)
        x_lo = torch.empty_like(x_hi)
        ops.stem_pack(x, perm, x_hi, x_lo)
        w_hi, w_lo = prep_weight(w, kind=1)
        M = N * P * Q
        out = torch.full((M, 64), float("nan"), device=DEV)
        geom = dict(sg["geom"], batch=N)
        ops.conv_fwd(x_hi, x_lo, w_hi, w_lo, out, M, 64, 256, passes=3, geom=geom)
        torch.cuda.synchronize()
        ref = F.conv2d(x[perm].double(), w.double(), stride=2, padding=3).permute(0, 2, 3, 1).reshape(M, 64)
        report("stem 7x7/2 %dx%d" % (H, W), rel(out, ref), 3e-5)


class _BN:
    def __init__(self, C, g):
        self.weight = (torch.rand((C,), generator=g) + 0.5).to(DEV)
        self.bias = torch.randn((C,), generator=g).to(DEV)
        self.running_mean = torch.randn((C,), generator=g).to(DEV)
        self.running_var = (torch.rand((C,), generator=g) + 0.5).to(DEV)
        self.num_batches_tracked = torch.zeros((), dtype=torch.int64, device=DEV)


def bn_coef_ref(raw, bn, train, rm, rv):
    """fp64 reference of (scale, shift, new running mean, new running var)"""
    raw = raw.double()
    if train:
        mean, var = raw.mean(0), raw.var(0, unbiased=False)
        n = raw.shape[0]
        rm_new = 0.9 * rm.double() + 0.1 * mean
        rv_new = 0.9 * rv.double() + 0.1 * var * n / (n - 1)
    else:
        mean, var, rm_new, rv_new = rm.double(), rv.double(), rm.double(), rv.double()
    sc = bn.weight.double() / torch.sqrt(var + 1e-5)
    return sc, bn.bias.double() - mean * sc, rm_new, rv_new


def make_coef(raw, bn, train):
    """coef tensor as the kernels expect it ([scale | shift], fp32), computed in torch for the apply-only checks"""
    sc, sh, _, _ = bn_coef_ref(raw, bn, train, bn.running_mean, bn.running_var)
    return torch.cat((sc, sh)).float().contiguous()


def bn_ref(raw, bn, train, rm, rv):
    raw = raw.double()
    if train:
        mean, var = raw.mean(0), raw.var(0, unbiased=False)
        n = raw.shape[0]
        rm_new = 0.9 * rm.double() + 0.1 * mean
        rv_new = 0.9 * rv.double() + 0.1 * var * n / (n - 1)
    else:
        mean, var, rm_new, rv_new = rm.double(), rv.double(), rm.double(), rv.double()
    y = (raw - mean) / torch.sqrt(var + 1e-5) * bn.weight.double() + bn.bias.double()
    return y, rm_new, rv_new


def check_bn_apply():
    g = torch.Generator().manual_seed(9)
    for C in (64, 256, 2048):
        M = 1000
        raw = (torch.randn((M, C), generator=g) * 2 + 0.5).to(DEV)
        raw2 = torch.randn((M, C), generator=g).to(DEV)
        res = torch.randn((M, C), generator=g).to(DEV)
        for train in (True, False):
            for res_kind in (0, 1, 2):
                bn, bn2 = _BN(C, g), _BN(C, g)
                coef, coef2 = make_coef(raw, bn, train), make_coef(raw2, bn2, train)
                out_hi = torch.empty((M, C), device=DEV, dtype=torch.float16)
                out_lo = torch.empty_like(out_hi)
                out_f = torch.empty((M, C), device=DEV)
                kw = {}
                r_hi, r_lo = split(res)
                if res_kind == 1:
                    kw["res_planes"] = (r_hi, r_lo)
                elif res_kind == 2:
                    kw["res_bn"] = ops.bn_side(raw2, coef2)
                ops.bn_apply(ops.bn_side(raw, coef), M, C, True, out_hi, out_lo, out_f, **kw)
                torch.cuda.synchronize()
                y, _, _ = bn_ref(raw, bn, train, bn.running_mean, bn.running_var)
                if res_kind == 1:
                    y = y + (r_hi.double() + r_lo.double())
                elif res_kind == 2:
                    y2, _, _ = bn_ref(raw2, bn2, train, bn2.running_mean, bn2.running_var)
                    y = y + y2
                y = F.relu(y)
                tag = "bn_apply C%d %s res%d" % (C, "train" if train else "eval", res_kind)
                report(tag + " f32", rel(out_f, y), 2e-6)
                report(tag + " hi+lo", rel(out_hi.double() + out_lo.double(), y), 2e-5)
        # eval-mode coefficient kernel
        bn = _BN(C, g)
        coef = torch.zeros((2 * C,), device=DEV)
        ops.build_bn_eval_coef(bn, coef)()
        torch.cuda.synchronize()
        sc, sh, _, _ = bn_coef_ref(raw, bn, False, bn.running_mean, bn.running_var)
        report("bn_eval_coef C%d" % C, rel(coef[:C], sc) + rel(coef[C:], sh), 1e-6)


def check_maxpool_finalpool():
    g = torch.Generator().manual_seed(10)
    N, P, Q, C = 3, 16, 12, 64
    raw = torch.randn((N, P, Q, C), generator=g).to(DEV)
    bn = _BN(C, g)
    flat = raw.reshape(-1, C)
    P2, Q2 = (P - 1) // 2 + 1, (Q - 1) // 2 + 1
    out_hi = torch.empty((N, P2, Q2, C), device=DEV, dtype=torch.float16)
    out_lo = torch.empty_like(out_hi)
    ops.bn_relu_maxpool(ops.bn_side(raw, make_coef(flat, bn, True)), out_hi, out_lo, N, P, Q, C)
    torch.cuda.synchronize()
    y, _, _ = bn_ref(flat, bn, True, bn.running_mean, bn.running_var)
    y = F.relu(y).reshape(N, P, Q, C).permute(0, 3, 1, 2)
    ref = F.max_pool2d(y, 3, 2, 1).permute(0, 2, 3, 1)
    report("bn_relu_maxpool", rel(out_hi.double() + out_lo.double(), ref), 2e-5)
    N, P, Q, C = 2, 112, 112, 64
    raw = torch.randn((N, P, Q, C), generator=g).to(DEV)
    flat = raw.reshape(-1, C)
    out_hi = torch.empty((N, 56, 56, C), device=DEV, dtype=torch.float16)
    out_lo = torch.empty_like(out_hi)
    ops.bn_relu_maxpool(ops.bn_side(raw, make_coef(flat, bn, True)), out_hi, out_lo, N, P, Q, C)
    torch.cuda.synchronize()
    y, _, _ = bn_ref(flat, bn, True, bn.running_mean, bn.running_var)
    ref = F.max_pool2d(F.relu(y).reshape(N, P, Q, C).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    report("bn_relu_maxpool 112x112", rel(out_hi.double() + out_lo.double(), ref), 2e-5)
    # final pool
    for C in (512, 2048):
        N, HW = 5, 49
        raw = torch.randn((N * HW, C), generator=g).to(DEV)
        res = torch.randn((N * HW, C), generator=g).to(DEV)
        r_hi, r_lo = split(res)
        bn = _BN(C, g)
        perm = torch.randperm(N, generator=g).to(DEV)
        spatial = torch.empty((N, C, 7, 7), device=DEV)
        pooled = torch.empty((N, C), device=DEV)
        ops.bn_final_pool(ops.bn_side(raw, make_coef(raw, bn, True)), N, HW, C, spatial, pooled, scatter_idx=perm,
                          res_planes=(r_hi, r_lo))
        torch.cuda.synchronize()
        y, _, _ = bn_ref(raw, bn, True, bn.running_mean, bn.running_var)
        y = F.relu(y + r_hi.double() + r_lo.double()).reshape(N, HW, C).permute(0, 2, 1).reshape(N, C, 7, 7)
        ref_sp = torch.empty_like(y)
        ref_sp[perm] = y
        report("bn_final_pool C%d spatial" % C, rel(spatial, ref_sp), 2e-6)
        report("bn_final_pool C%d pooled" % C, rel(pooled, ref_sp.mean((2, 3))), 2e-6)


def check_small():
    g = torch.Generator().manual_seed(12)
    x = torch.randn((37, 128), generator=g).to(DEV)
    out = torch.empty_like(x)
    ops.l2_normalize(x, out)
    report("l2_normalize", rel(out, F.normalize(x.double(), dim=1)), 1e-6)
    hi = torch.empty(x.shape, device=DEV, dtype=torch.float16)
    lo = torch.empty_like(hi)
    ops.split_f16(x, hi, lo)
    report("split_f16", rel(hi.double() + lo.double(), x), 2e-5)
    r = torch.empty_like(x)
    ops.round_tf32(x, r)
    report("round_tf32", rel(r, x), 3e-4, "low bits zero: %s" % bool(((r.view(torch.int32) & 0x1FFF) == 0).all()))
    xi = torch.randn((3, 3, 50, 47), generator=g).to(DEV)
    perm = torch.randperm(3, generator=g).to(DEV)
    out = torch.empty((27, 3, 17, 16), device=DEV)
    ops.jigsaw_patchify(xi, perm, out)
    report("jigsaw_patchify", rel(out, vo.jigsaw_patchify(xi[perm].cpu())), 0)
    feats = torch.randn((18, 64), generator=g).to(DEV)
    order = torch.stack([torch.randperm(9, generator=g) for _ in range(2)]).to(DEV)
    outg = torch.empty((2, 9 * 64), device=DEV)
    ops.jigsaw_gather(feats, order, outg)
    f = feats.reshape(2, 9, 64)
    refg = f[torch.arange(2, device=DEV)[:, None].expand(-1, 9), order].reshape(2, -1)
    report("jigsaw_gather", rel(outg, refg), 0)


def ema_table(pairs):
    chunks = []
    for dst, src in pairs:
        n = dst.numel()
        for off in range(0, n, 8192):
            chunks.append((dst.data_ptr() + 4 * off, src.data_ptr() + 4 * off, min(8192, n - off)))
    arr = np.array(chunks, dtype=np.int64).reshape(-1, 3)
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(DEV), len(chunks)


def check_ema_enqueue():
    g = torch.Generator().manual_seed(13)
    shapes = [(64, 3, 7, 7), (64,), (128, 64, 3, 3), (1000, 512), (1000,), (7,), (20001,)]
    keyp = [torch.randn(s, generator=g).to(DEV) for s in shapes]
    qp = [torch.randn(s, generator=g).to(DEV) for s in shapes]
    ref = [k.cpu().clone() for k in keyp]
    vo.param_update({i: r for i, r in enumerate(ref)}, {i: q.cpu() for i, q in enumerate(qp)}, 0.999, range(len(ref)))
    tab, n = ema_table(list(zip(keyp, qp)))
    K, D = 40, 16
    queue = torch.randn((K, D), generator=g).to(DEV)
    qt = torch.zeros_like(queue)
    oq = vo.StorageQueue(K, D, init=queue.cpu())
    tail = 0
    ok_bits = True
    for step, nk in enumerate([16, 16, 16, 8, 40]):
        keys = torch.randn((nk, D), generator=g).to(DEV)
        tail, wrapped = ops.ema_enqueue(tab if step == 0 else None, n if step == 0 else 0, 0.999, queue, qt, keys, tail)
        oq.enqueue(keys.cpu())
        torch.cuda.synchronize()
        ok_bits &= torch.equal(queue.cpu(), oq.vector_queue) and tail == oq.current_tail
    report("enqueue ring buffer bit-exact", 0.0 if ok_bits else 1.0, 0)
    rq = torch.empty_like(queue)
    ops.round_tf32(queue, rq)
    written = qt != 0
    report("enqueue tf32 shadow", rel(qt[written], rq[written]), 0)
    worst = max(rel(k, r) for k, r in zip(keyp, ref))
    exact = all(torch.equal(k.cpu(), r) for k, r in zip(keyp, ref))
    report("ema multi-tensor", worst, 1e-7, "bit-exact=%s" % exact)


def check_infonce():
    g = torch.Generator().manual_seed(21)
    cases = [("cfg0 B8 K1024 nf2", 8, 1024, 128, 2, 0.07), ("B256 K65536 nf4 T.07", 256, 65536, 128, 4, 0.07),
             ("B256 K4096 nf4 T.2", 256, 4096, 128, 4, 0.2), ("B128 K1000 D64 nf1", 128, 1000, 64, 1, 0.07),
             ("moco B64 K2048", 64, 2048, 128, 0, 0.07), ("self B32 K0 nf4", 32, 0, 128, 4, 0.03),
             ("B6 K100 D32 nf2", 6, 100, 32, 2, 0.07)]
    for name, B, K, D, nf, T in cases:
        q = F.normalize(torch.randn((B, D), generator=g), dim=1)
        k = F.normalize(q + 0.7 * torch.randn((B, D), generator=g), dim=1)
        if name.startswith("self"):
            k = q.clone()
        queue = F.normalize(torch.randn((max(K, 1), D), generator=g), dim=1)[:K]
        qd, kd, qud = q.to(DEV), k.to(DEV), queue.to(DEV)
        qt = torch.empty_like(qud)
        if K:
            ops.round_tf32(qud, qt)
        out = ops.infonce_fwd(qd, kd, qt if K else None, nf, T)
        torch.cuda.synchronize()
        if name.startswith("self"):
            sims = q.double() @ q.double().t()
            mask = vo.block_diag_mask(B, nf, B)
        else:
            fw = vo.vince_forward(q.double(), k.double(), queue.double(), max(nf, 1), inter_batch_comparison=nf > 0)
            sims, mask = fw["vince_similarities"], fw["vince_similarities_mask"]
        ce = vo.similarity_cross_entropy(sims, T, mask)
        mt = vo.get_metrics(sims, mask, ce["softmax_weight"])
        sc = out["scalars"].cpu().double()
        report("infonce %s loss" % name, abs(sc[0].item() - ce["dist"].item()) / abs(ce["dist"].item()), 1e-4,
               "(%.6f vs %.6f)" % (sc[0].item(), ce["dist"].item()))
        report("  dists", rel(out["dists"], ce["dists"].reshape(B, -1)), 1e-4)
        report("  weights", rel(out["weights"], ce["softmax_weights"].reshape(B, -1)), 1e-3)
        report("  softmax_weight", abs(sc[1].item() - ce["softmax_weight"].item()) / max(ce["softmax_weight"].item(), 1e-12), 1e-3)
        report("  accuracy", abs(sc[2].item() - mt["nce_accuracy_mean"].item()), 1e-2)
        report("  cosine_sim", abs(sc[3].item() - mt["cosine_sim"].item()), 1e-5)
        if K or nf > 0 and B > nf:
            report("  cosine_sim_neg_max", abs(sc[4].item() - mt["cosine_sim_neg_max"].item()), 1e-4)


CHECKS = [check_small, check_ema_enqueue, check_bn_apply, check_maxpool_finalpool, check_gemm_tiled, check_conv_im2col,
          check_conv_halo, check_stem, check_infonce]


def main():
    filt = sys.argv[1:]
    print("device:", torch.cuda.get_device_name(0))
    for fn in CHECKS:
        if filt and not any(f in fn.__name__ for f in filt):
            continue
        print("==", fn.__name__, flush=True)
        try:
            fn()
        except Exception:
            RESULTS.append((fn.__name__, False))
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:  # sticky CUDA error: nothing else can run in this process
                print("CUDA context is dead:", e)
                break
    bad = [n for n, ok in RESULTS if not ok]
    print("SUMMARY: %d checks, %d failed" % (len(RESULTS), len(bad)))
    for n in bad:
        print("  FAILED:", n)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
