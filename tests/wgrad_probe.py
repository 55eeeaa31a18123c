"""GPU diagnostic: the batched split-K GEMM mode of vince_conv_fwd against torch (not collected by pytest)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vince_b200 import ops  # noqa: E402

dev = "cuda"


def split(x):
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    return hi.contiguous(), lo.contiguous()


def case(name, M, N, K, kchunk, taps, shift_w):
    g = torch.Generator().manual_seed(0)
    A = torch.randn((M, K), generator=g).to(dev)
    B = torch.randn((N, K), generator=g).to(dev)
    a_hi, a_lo = split(A)
    if taps == 9:       # three column-shifted copies stacked along the rows: copy_j[k] = B[k + j - 1]
        Bz = torch.zeros((3, N, K), device=dev)
        Bz[0, :, 1:] = B[:, :-1]
        Bz[1] = B
        Bz[2, :, :-1] = B[:, 1:]
        b_hi, b_lo = split(Bz.reshape(3 * N, K))
    else:
        b_hi, b_lo = split(B)
    mpad = (M + 127) // 128 * 128
    splits = (K + kchunk - 1) // kchunk
    out = torch.full((taps * splits * mpad, N), float("nan"), device=dev)
    try:
        ops.conv_fwd(a_hi, a_lo, b_hi, b_lo, out, M, N, K, passes=3, kchunk=kchunk, taps=taps, shift_w=shift_w)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("%-40s FAILED: %s" % (name, str(e)[:150]))
        return
    worst = 0.0
    Ad, Bd = A.double(), B.double()
    for tap in range(taps):
        sh = ((tap // 3 - 1) * shift_w + (tap % 3 - 1)) if taps == 9 else 0
        Bs = torch.zeros_like(Bd)
        if sh >= 0:
            Bs[:, :K - sh] = Bd[:, sh:]
        else:
            Bs[:, -sh:] = Bd[:, :K + sh]
        ref = Ad @ Bs.t()
        got = out.view(taps, splits, mpad, N)[tap, :, :M].double().sum(0)
        worst = max(worst, ((got - ref).norm() / ref.norm()).item())
    print("%-40s rel err %.2e" % (name, worst))


case("plain batched path", 64, 64, 256, 256, 1, 0)
case("split-K 4", 64, 64, 1024, 256, 1, 0)
case("split-K ragged", 192, 128, 1000, 192, 1, 0)
case("9 taps", 64, 64, 512, 512, 9, 8)
case("9 taps + splits", 512, 512, 256, 128, 9, 8)
case("9 taps big", 128, 64, 8 * 58 * 64, 4096, 9, 64)
