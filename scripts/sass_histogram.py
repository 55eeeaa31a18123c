"""SASS opcode evidence for libvince_b200.so (no GPU needed):  python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt
Per kernel family: number of instantiations and counts of the mnemonics that prove a Blackwell-native path
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit, LDGSTS = cp.async,
SYNCS = mbarrier) next to the legacy tensor path (HMMA = mma.sync, should be 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vince_b200", "csrc", "libvince_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP",
         "LDGSTS", "SYNCS", "HMMA", "HGMMA", "DADD", "DFMA", "BAR.SYNC", "FENCE.VIEW.ASYNC", "ATOMG", "RED", "STG", "LDG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fam = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            base = re.sub(r"<.*", "", name.split("(")[0]).replace("void ", "")
            cur = fam.setdefault(base, dict(n=0, instr=0, ops=collections.Counter()))
            cur["n"] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["instr"] += 1
            for w in WATCH:
                if op.startswith(w):
                    cur["ops"][w] += 1
    print("SASS opcode histogram of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(LIB, ROOT))
    print("%-34s %5s %8s  %s" % ("kernel family", "inst.", "SASS", "watched mnemonics"))
    tot = collections.Counter()
    for base, d in fam.items():
        ops = "  ".join("%s=%d" % (k, v) for k, v in sorted(d["ops"].items()))
        print("%-34s %5d %8d  %s" % (base[:34], d["n"], d["instr"], ops))
        tot.update(d["ops"])
    print("TOTAL: " + "  ".join("%s=%d" % (k, v) for k, v in sorted(tot.items())))
    print("legacy tensor path (HMMA / HGMMA): %d" % (tot["HMMA"] + tot["HGMMA"]))


if __name__ == "__main__":
    sys.exit(main())
