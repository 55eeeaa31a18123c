"""Build recipe for libvince_b200.so (hand-written sm_100a CUDA behind a C ABI).

    python -m vince_b200.build            # incremental in-tree build
    python -m vince_b200.build --force    # rebuild everything

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE (vince_b200/csrc/libvince_b200.so,
git-ignored) so it travels with the source snapshot to the GPU box; cudart is linked statically and the driver
API / NCCL are bound at run time, so the .so loads on a GPU-less machine too.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libvince_b200.so")
SOURCES = ["host.cu", "conv_gemm.cu", "elementwise.cu", "infonce.cu", "infonce_bwd.cu", "knn.cu", "backward.cu", "capi.cu"]
HEADERS = ["common.cuh", "kernels.h", os.path.join("..", "..", "include", "vince_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, force):
    obj = os.path.join(CSRC, src.replace(".cu", ".o"))
    stamp = obj + ".sha"
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    dig = _digest(deps)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, r.stderr


def build(force=False, verbose=False):
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, force), SOURCES))
    objs = [o for o, _ in results]
    rebuilt = any(log for _, log in results)
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print("built", path)
