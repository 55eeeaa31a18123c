"""Device time of the fused InfoNCE forward alone (GPU box only; not collected by pytest): 20 launches captured in a CUDA
graph, replayed and timed with CUDA events - no host issue time, no profiler."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vince_b200 import ops  # noqa: E402

dev = "cuda:0"
for B, K, D, nf, T in ((256, 65536, 128, 4, 0.07), (256, 65536, 128, 4, 0.2), (128, 131072, 128, 4, 0.2)):
    q = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1)
    k = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=1)
    queue = torch.nn.functional.normalize(torch.randn(K, D, device=dev), dim=1)
    qt = torch.empty_like(queue)
    ops.round_tf32(queue, qt)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(3):
        out = ops.infonce_fwd(q, k, qt, nf, T)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    n = 20
    with torch.cuda.stream(s):
        g.capture_begin()
        for _ in range(n):
            out = ops.infonce_fwd(q, k, qt, nf, T)
        g.capture_end()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    bytes_ = 4.0 * D * (K + 2 * B)
    t = sorted(ts)[len(ts) // 2]
    print("infonce_fwd B=%d K=%d D=%d T=%.2f: %.1f us per launch (memset + kernel, back to back, queue L2-warm after the "
          "first), %.0f GB/s algorithmic, loss %.5f" % (B, K, D, T, t, bytes_ / t / 1e3, float(out["scalars"][0])))
