"""Debug timeline of the conv kernel's three roles for CTA 0 (needs the -DVB_TRACE build of the library):
    VINCE_B200_LIB=vince_b200/csrc/libvince_b200_trace.so python tests/conv_trace.py --shape layer3 [--halo 0]
tags: 0 producer got a free B slot, 1 MMA warp saw data, 2 MMA warp finished issuing a k-block, 3 epilogue saw a tile."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vince_b200 import ops  # noqa: E402
from conv_bench import SHAPES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="r18.layer3.3x3")
    ap.add_argument("--halo", type=int, default=0)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    name, H, W, Cin, Cout, R, stride, pad = [s for s in SHAPES if a.shape in s[0]][0]
    dev = "cuda"
    B = a.batch
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    M, K = B * P * Q, R * R * Cin
    hi = torch.randn((B, H, W, Cin), device=dev).to(torch.float16)
    lo = (torch.randn((B, H, W, Cin), device=dev) * 0.01).to(torch.float16)
    w_hi = (torch.randn((Cout, K), device=dev) * 0.05).to(torch.float16)
    w_lo = (torch.randn((Cout, K), device=dev) * 0.0005).to(torch.float16)
    out = torch.empty((M, Cout), device=dev)
    stats = torch.zeros((2 * Cout,), device=dev, dtype=torch.float64)
    geom = None
    if not (R == 1 and stride == 1):
        geom = dict(batch=B, H=H, W=W, Cin=Cin, R=R, S=R, stride=stride, pad_lo_h=pad, pad_lo_w=pad, pad_hi_h=pad, pad_hi_w=pad)
    x_hi = hi.reshape(-1, Cin) if geom is None else hi
    x_lo = lo.reshape(-1, Cin) if geom is None else lo
    run = ops.build_conv_fwd(x_hi, x_lo, w_hi, w_lo, out, M, Cout, K, geom=geom, block_n=a.bn, stats=stats, halo_mode=a.halo)
    run()
    torch.cuda.synchronize()
    trace = torch.zeros((16, 512), device=dev, dtype=torch.int64)
    os.environ["VINCE_B200_TRACE_PTR"] = str(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    del os.environ["VINCE_B200_TRACE_PTR"]
    full = trace.cpu().numpy()
    t = full[:8]
    t0 = full[full > 0].min()
    if (full[8:] > 0).any():
        print("pair trace (ns since first event): kb | leader: A-slot B-slot data issued | peer: A-slot B-slot data-local")
        for i in range(24):
            print("%3d | %7d %7d %7d %7d | %7d %7d %7d" % (i, full[4][i] - t0, full[0][i] - t0, full[1][i] - t0, full[2][i] - t0,
                                                           full[12][i] - t0, full[8][i] - t0, full[13][i] - t0))
    n = int((t[1] > 0).sum())
    print("%s: %d k-blocks traced for CTA 0 (cycles relative to first event)" % (name, n))
    print("  kb   prod_slot   mma_data  mma_issued   d(data)  d(issued)  issue_len   data-prod")
    for i in range(min(n, 80)):
        p_, m1, m2 = t[0][i] - t0, t[1][i] - t0, t[2][i] - t0
        d1 = t[1][i] - t[1][i - 1] if i else 0
        d2 = t[2][i] - t[2][i - 1] if i else 0
        print("%4d %11d %10d %11d %9d %10d %10d %11d" % (i, p_, m1, m2, d1, d2, m2 - m1, m1 - p_))
    d = np.diff(t[1][:n])
    print("steady-state cycles per k-block: median %d mean %.0f" % (np.median(d), d.mean()))
    ep = t[3][t[3] > 0] - t0
    print("epilogue tile-ready times:", ep[:8].tolist())


if __name__ == "__main__":
    main()
