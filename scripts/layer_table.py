"""Per-layer table from an ncu launch list (gpu__time_duration.sum) of one bench step: maps the conv_gemm launches of
one encoder forward, in plan order, to their implicit-GEMM shapes and prints time vs the tensor (3 fp16 passes) and
HBM bounds.   python scripts/layer_table.py gpurun_out/launches_cfg2.csv ResNet50 256 224"""
import collections
import csv
import sys


def conv_shapes(backbone, B, H):
    kind, layers = {"ResNet18": ("basic", (2, 2, 2, 2)), "ResNet50": ("bottleneck", (3, 4, 6, 3))}[backbone]
    out = []
    P = (H - 1) // 2 + 1
    out.append(("stem", B * P * P, 64, 147, 256, 7, 2))
    h = (P - 1) // 2 + 1
    inpl = 64
    exp = 1 if kind == "basic" else 4
    for li, nb in enumerate(layers):
        planes = 64 * 2 ** li
        for b in range(nb):
            stride = 2 if (b == 0 and li > 0) else 1
            ho = (h - 1) // stride + 1
            name = "l%d.%d" % (li + 1, b)
            if kind == "basic":
                out.append((name + ".c1 3x3/%d" % stride, B * ho * ho, planes, 9 * inpl, 9 * inpl, 3, stride))
                out.append((name + ".c2 3x3", B * ho * ho, planes, 9 * planes, 9 * planes, 3, 1))
            else:
                out.append((name + ".c1 1x1", B * h * h, planes, inpl, inpl, 1, 1))
                out.append((name + ".c2 3x3/%d" % stride, B * ho * ho, planes, 9 * planes, 9 * planes, 3, stride))
                out.append((name + ".c3 1x1", B * ho * ho, planes * exp, planes, planes, 1, 1))
            if b == 0 and (stride != 1 or inpl != planes * exp):
                out.append((name + ".ds 1x1/%d" % stride, B * ho * ho, planes * exp, inpl, inpl, 1, stride))
            inpl = planes * exp
            h = ho
    return out


def main():
    path, backbone, B, H = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    g = hdr.index("Grid Size")
    convs = [(r[k].split("(")[0].replace("void vb::", ""), float(r[v].replace(",", "")) / 1e3, r[g]) for r in rows[hi + 2:]
             if "conv_gemm" in r[k]]
    shapes = conv_shapes(backbone, B, H)
    n = len(shapes)
    print("%d conv launches in the list, %d convs per encoder forward" % (len(convs), n))
    tot = collections.Counter()
    print("%-18s %9s %5s %5s  %-28s %8s %8s %8s %6s" % ("layer", "M", "N", "K", "kernel", "us", "tens_us", "hbm_us", "eff"))
    for i, (name, M, N, K, Kp, R, stride) in enumerate(shapes):
        kern, us, grid = convs[i]
        flops = 2.0 * M * N * K
        tens = 3 * 2.0 * M * N * Kp / 1.6777e15 * 1e6          # 3 passes at the measured burst bf16 rate
        in_elems = M * (Kp if R == 1 and stride == 1 else (K // (R * R)) * stride * stride)
        hbm = (4.0 * in_elems + 4.0 * M * N) / 6.553e12 * 1e6
        bound = max(tens, hbm)
        print("%-18s %9d %5d %5d  %-28s %8.1f %8.1f %8.1f %6.2f" % (name, M, N, K, kern + " g" + grid.strip("()").split(",")[0], us, tens, hbm, bound / us))
        tot["us"] += us
        tot["bound"] += bound
        tot["flops"] += flops
    print("total %.1f us, sum of bounds %.1f us (%.2f), %.1f algorithmic TFLOP/s" % (tot["us"], tot["bound"], tot["bound"] / tot["us"], tot["flops"] / tot["us"] / 1e6))


if __name__ == "__main__":
    main()
