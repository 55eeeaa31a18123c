"""Micro-benchmark of the HBM-bound encoder kernels at the benchmark shapes (GPU box only; not collected by pytest).

    python tests/elem_bench.py [--batch 256] [--iters 7]

Prints per-kernel time (CUDA events; three distinct buffer sets cycled so nothing is served from L2) and the
algorithmic GB/s (bytes each kernel must read + write once).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vince_b200 import ops  # noqa: E402


def timeit(runs, iters):
    for r in runs:
        r()
    torch.cuda.synchronize()
    times = []
    for it in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        runs[it % len(runs)]()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    return sorted(times)[len(times) // 2], min(times)


def report(name, t, tmin, nbytes, flops=None):
    extra = "" if flops is None else "  %7.1f TFLOP/s (alg)" % (flops / t / 1e6)
    print("%-44s %8.1f us  %7.0f GB/s  (min %.1f us, %.0f MB)%s" % (name, t, nbytes / t / 1e3, tmin, nbytes / 1e6, extra),
          flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--only", default="", help="substring filter: stem / bn_apply")
    a = ap.parse_args()
    dev = "cuda"
    B, nbuf = a.batch, 3
    if not a.only or a.only in "stem":
        bench_stem(a, dev, B, nbuf)
    if not a.only or a.only in "bn_apply":
        bench_bn_apply(a, dev, B, nbuf)
    if not a.only or a.only in "nce":
        bench_nce(a, dev, B)


def bench_nce(a, dev, B):
    """fused InfoNCE forward / backward at K=65536, D=128 (three queue copies cycled: 3 x 32 MiB > L2 per pass is not
    needed here, the 32 MiB queue itself is what a step streams)."""
    K, D, nf, T = 65536, 128, 4, 0.07
    F = torch.nn.functional
    q = F.normalize(torch.randn((B, D), device=dev), dim=1)
    k = F.normalize(q + 0.5 * torch.randn((B, D), device=dev), dim=1)
    queues = []
    for _ in range(3):
        qu = F.normalize(torch.randn((K, D), device=dev), dim=1)
        qt = torch.empty_like(qu)
        ops.round_tf32(qu, qt)
        queues.append(qt)
    fwd = [ops.infonce_fwd(q, k, queues[i], nf, T) for i in range(3)]
    t, tm = timeit([lambda i=i: ops.infonce_fwd(q, k, queues[i], nf, T) for i in range(3)], a.iters)
    report("infonce_fwd B=%d K=%d D=%d (4 launches)" % (B, K, D), t, tm, 4 * D * (K + 2 * B), flops=2.0 * B * (B + K) * D)
    t, tm = timeit([lambda i=i: ops.infonce_bwd(q, k, queues[i], nf, T, fwd[i]) for i in range(3)], a.iters)
    report("infonce_bwd B=%d K=%d D=%d (4 launches)" % (B, K, D), t, tm, 4 * D * (K + 2 * B), flops=4.0 * B * (B + K) * D)


def bench_stem(a, dev, B, nbuf):
    H = W = 224
    sg = ops.stem_geometry(H, W)
    P, Q, Ha, Wb = sg["P"], sg["Q"], sg["Ha"], sg["Wb"]
    # ---- stem pack ----
    xs = [torch.randn((B, 3, H, W), device=dev) for _ in range(nbuf)]
    packs = [(torch.empty((B, Ha, Wb, 16), device=dev, dtype=torch.float16),
              torch.empty((B, Ha, Wb, 16), device=dev, dtype=torch.float16)) for _ in range(nbuf)]
    perm = torch.randperm(B, device=dev)
    runs = [ops.build_stem_pack(xs[i], perm, packs[i][0], packs[i][1]) for i in range(nbuf)]
    t, tm = timeit(runs, a.iters)
    report("stem_pack 224x224", t, tm, B * 3 * H * W * 4 + 2 * packs[0][0].numel() * 2)
    # ---- stem conv ----
    w_hi = (torch.randn((64, 256), device=dev) * 0.05).to(torch.float16)
    w_lo = (torch.randn((64, 256), device=dev) * 0.0005).to(torch.float16)
    M = B * P * Q
    raws = [torch.empty((M, 64), device=dev) for _ in range(nbuf)]
    stats = torch.zeros((128,), device=dev, dtype=torch.float64)
    runs = [ops.build_conv_fwd(packs[i][0], packs[i][1], w_hi, w_lo, raws[i], M, 64, 256, passes=3,
                               geom=dict(sg["geom"], batch=B), stats=stats) for i in range(nbuf)]
    t, tm = timeit(runs, a.iters)
    report("stem conv (M=%d N=64 K=256)" % M, t, tm, 2 * packs[0][0].numel() * 2 + M * 64 * 4, flops=2.0 * M * 64 * 147)
    # ---- bn + relu + maxpool ----
    coef = torch.cat([torch.rand((64,), device=dev) + 0.5, torch.randn((64,), device=dev)])
    P2, Q2 = (P - 1) // 2 + 1, (Q - 1) // 2 + 1
    pooled = [(torch.empty((B * P2 * Q2, 64), device=dev, dtype=torch.float16),
               torch.empty((B * P2 * Q2, 64), device=dev, dtype=torch.float16)) for _ in range(nbuf)]
    runs = [ops.build_bn_relu_maxpool(ops.bn_side(raws[i], coef), pooled[i][0], pooled[i][1], B, P, Q, 64)
            for i in range(nbuf)]
    t, tm = timeit(runs, a.iters)
    report("bn_relu_maxpool 112x112x64", t, tm, M * 64 * 4 + 2 * pooled[0][0].numel() * 2)


def bench_bn_apply(a, dev, B, nbuf):
    # ---- bn_apply at the four ResNet-18 stage shapes, without / with residual planes ----
    for (hw, C) in [(56, 64), (28, 128), (14, 256), (7, 512), (56, 256)]:
        M = B * hw * hw
        coef = torch.cat([torch.rand((C,), device=dev) + 0.5, torch.randn((C,), device=dev)])
        raw = [torch.randn((M, C), device=dev) for _ in range(nbuf)]
        res = [(torch.randn((M, C), device=dev).to(torch.float16), torch.randn((M, C), device=dev).to(torch.float16))
               for _ in range(nbuf)]
        out = [(torch.empty((M, C), device=dev, dtype=torch.float16), torch.empty((M, C), device=dev, dtype=torch.float16))
               for _ in range(nbuf)]
        runs = [ops.build_bn_apply(ops.bn_side(raw[i], coef), M, C, True, out[i][0], out[i][1]) for i in range(nbuf)]
        t, tm = timeit(runs, a.iters)
        report("bn_apply %dx%dx%d" % (hw, hw, C), t, tm, M * C * 8)
        runs = [ops.build_bn_apply(ops.bn_side(raw[i], coef), M, C, True, out[i][0], out[i][1], res_planes=res[i])
                for i in range(nbuf)]
        t, tm = timeit(runs, a.iters)
        report("bn_apply %dx%dx%d + residual" % (hw, hw, C), t, tm, M * C * 12)
        del raw, res, out


if __name__ == "__main__":
    main()
