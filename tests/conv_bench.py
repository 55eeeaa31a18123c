"""Micro-benchmark of the tcgen05 conv kernel on the ResNet-18/50 layer shapes at the benchmark batch (GPU box only).

    python tests/conv_bench.py [--batch 256] [--filter layer1] [--halo -1|0|1] [--bn 0|64|128|256] [--iters 5]

Prints per-shape time (CUDA events, L2 flushed between iterations by cycling through distinct buffers) and the
algorithmic TFLOP/s (2*M*N*K, no credit for the three fp16 passes).  Not collected by pytest.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vince_b200 import ops  # noqa: E402

# name, H, W, Cin, Cout, R, stride, pad
SHAPES = [
    ("r18.layer1.3x3", 56, 56, 64, 64, 3, 1, 1),
    ("r18.layer2.0.conv1 3x3/2", 56, 56, 64, 128, 3, 2, 1),
    ("r18.layer2.3x3", 28, 28, 128, 128, 3, 1, 1),
    ("r18.layer2.ds 1x1/2", 56, 56, 64, 128, 1, 2, 0),
    ("r18.layer3.0.conv1 3x3/2", 28, 28, 128, 256, 3, 2, 1),
    ("r18.layer3.3x3", 14, 14, 256, 256, 3, 1, 1),
    ("r18.layer4.0.conv1 3x3/2", 14, 14, 256, 512, 3, 2, 1),
    ("r18.layer4.3x3", 7, 7, 512, 512, 3, 1, 1),
    ("r18.layer4.ds 1x1/2", 14, 14, 256, 512, 1, 2, 0),
    ("r50.layer1 1x1 64->256", 56, 56, 64, 256, 1, 1, 0),
    ("r50.layer1 1x1 256->64", 56, 56, 256, 64, 1, 1, 0),
    ("r50.layer2 1x1 128->512", 28, 28, 128, 512, 1, 1, 0),
    ("r50.layer2 1x1 512->128", 28, 28, 512, 128, 1, 1, 0),
    ("r50.layer3 1x1 1024->256", 14, 14, 1024, 256, 1, 1, 0),
    ("r50.layer4 1x1 512->2048", 7, 7, 512, 2048, 1, 1, 0),
    ("r50.layer3 1x1 256->1024", 14, 14, 256, 1024, 1, 1, 0),
    ("r50.layer2 1x1 256->128", 56, 56, 256, 128, 1, 1, 0),
    ("r50.layer4 1x1 2048->512", 7, 7, 2048, 512, 1, 1, 0),
    ("r50.layer3 3x3", 14, 14, 256, 256, 3, 1, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--filter", default="")
    ap.add_argument("--halo", type=int, default=-1)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--nostats", action="store_true")
    ap.add_argument("--statsonly", action="store_true")
    ap.add_argument("--tstats", action="store_true", help="transposed statistics pass (1x1 stride-1 shapes only)")
    ap.add_argument("--apply", type=int, default=-1, help="apply epilogue: 0 no residual, 1 residual planes, 2 bn(raw) residual")
    a = ap.parse_args()
    dev = "cuda"
    B = a.batch
    if not a.filter or a.filter in "stem 7x7/2":
        # the stem runs over the packed overlapping-window layout (ops.stem_geometry): 224x224 frames -> 112x112x64
        sg = ops.stem_geometry(224, 224)
        M = B * sg["P"] * sg["Q"]
        xs = [(torch.randn((B, sg["Ha"], sg["Wb"], 16), device=dev).to(torch.float16),
               (torch.randn((B, sg["Ha"], sg["Wb"], 16), device=dev) * 0.01).to(torch.float16)) for _ in range(3)]
        w_hi = (torch.randn((64, 256), device=dev) * 0.05).to(torch.float16)
        w_lo = (torch.randn((64, 256), device=dev) * 0.0005).to(torch.float16)
        outs = [torch.empty((M, 64), device=dev) for _ in range(3)]
        stats = torch.zeros((128,), device=dev, dtype=torch.float64)
        runs = [ops.build_conv_fwd(xs[i][0], xs[i][1], w_hi, w_lo, outs[i], M, 64, 256, passes=3,
                                   geom=dict(sg["geom"], batch=B), stats=None if a.nostats else stats) for i in range(3)]
        for r in runs:
            r()
        torch.cuda.synchronize()
        times = []
        for it in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            runs[it % 3]()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(times)[len(times) // 2]
        print("%-28s M=%8d N=%4d K=%5d  %8.1f us  %7.1f TFLOP/s (alg)  min %.1f us" % ("stem 7x7/2 (packed K=256)", M, 64, 147,
              t, 2.0 * M * 64 * 147 / t / 1e6, min(times)), flush=True)
        del xs, outs
    for name, H, W, Cin, Cout, R, stride, pad in SHAPES:
        if a.filter and a.filter not in name:
            continue
        P = (H + 2 * pad - R) // stride + 1
        Q = (W + 2 * pad - R) // stride + 1
        M = B * P * Q
        K = R * R * Cin
        nbuf = 3
        acts = [(torch.randn((B, H, W, Cin), device=dev).to(torch.float16), torch.randn((B, H, W, Cin), device=dev).to(torch.float16) * 0.01)
                for _ in range(nbuf)]
        w_hi = (torch.randn((Cout, K), device=dev) * 0.05).to(torch.float16)
        w_lo = (torch.randn((Cout, K), device=dev) * 0.0005).to(torch.float16)
        outs = [torch.empty((M, Cout), device=dev) for _ in range(nbuf)]
        stats = torch.zeros((2 * Cout,), device=dev, dtype=torch.float64)
        geom = None
        if not (R == 1 and stride == 1):
            geom = dict(batch=B, H=H, W=W, Cin=Cin, R=R, S=R, stride=stride, pad_lo_h=pad, pad_lo_w=pad, pad_hi_h=pad, pad_hi_w=pad)
        runs = []
        if a.tstats and geom is not None:
            continue
        coef = torch.cat([torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev)])
        planes = [(torch.empty((M, Cout), device=dev, dtype=torch.float16), torch.empty((M, Cout), device=dev, dtype=torch.float16))
                  for _ in range(nbuf)] if a.apply >= 0 else None
        res_p = (torch.randn((M, Cout), device=dev).to(torch.float16), (torch.randn((M, Cout), device=dev) * 0.01).to(torch.float16)) \
            if a.apply == 1 else None
        res_r = torch.randn((M, Cout), device=dev) if a.apply == 2 else None
        for i in range(nbuf):
            hi, lo = acts[i]
            x_hi = hi.reshape(-1, Cin) if geom is None else hi
            x_lo = lo.reshape(-1, Cin) if geom is None else lo
            if a.apply >= 0:
                kw = {}
                if a.apply == 1:
                    kw["res_planes"] = res_p
                elif a.apply == 2:
                    kw["res_raw"], kw["res_coef"] = res_r, coef
                runs.append(ops.build_conv_fwd(x_hi, x_lo, w_hi, w_lo, None, M, Cout, K, passes=a.passes, geom=geom,
                                               block_n=a.bn, halo_mode=a.halo, relu=True, out_planes=planes[i],
                                               ep_coef=coef, **kw))
                continue
            so = 2 if a.tstats else (1 if a.statsonly else 0)
            runs.append(ops.build_conv_fwd(x_hi, x_lo if a.passes == 3 else None, w_hi, w_lo if a.passes == 3 else None,
                                           None if so else outs[i], M, Cout, K, passes=a.passes, geom=geom,
                                           block_n=a.bn, stats=None if a.nostats else stats, halo_mode=a.halo,
                                           stats_only=so))
        for r in runs:
            r()
        torch.cuda.synchronize()
        times = []
        for it in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            runs[it % nbuf]()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(times)[len(times) // 2]
        flops = 2.0 * M * Cout * K
        print("%-28s M=%8d N=%4d K=%5d  %8.1f us  %7.1f TFLOP/s (alg)  min %.1f us" % (name, M, Cout, K, t, flops / t / 1e6, min(times)),
              flush=True)


if __name__ == "__main__":
    main()
