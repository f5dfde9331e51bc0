#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the tcgen05 / TMEM / TMA / mbarrier / legacy-MMA mnemonics in every kernel of
libfld_sm100.so (cuobjdump -sass).  UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, IDP = integer dot product (alignment blend), HMMA = legacy mma.sync (must be 0).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "face-landmark-detector_b200", "lib", "libfld_sm100.so")
PATS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "SYNCS", "UTCATOM", "IDP", "HMMA", "IMMA", "MUFU.EX2", "STG", "LDG", "LDS", "STS", "DADD", "DMUL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for p in PATS:
            if op.startswith(p):
                kernels[cur][p] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    tot = collections.Counter()
    print("# %s" % __doc__.strip().splitlines()[0])
    print("# library: face-landmark-detector_b200/lib/libfld_sm100.so (nvcc -gencode arch=compute_100a,code=sm_100a), %d kernels" % len(kernels))
    print("%-86s %7s  %s" % ("kernel", "instr", "  ".join("%s" % p for p in PATS if p not in ("STG", "LDG", "LDS", "STS", "DADD", "DMUL"))))
    for k, c in kernels.items():
        n = re.sub(r"\(anonymous namespace\)::", "", names[k])
        n = re.sub(r"\((bool|int|unsigned int)\)", "", n)
        n = re.sub(r"\(.*", "", n)
        cols = [p for p in PATS if p not in ("STG", "LDG", "LDS", "STS", "DADD", "DMUL")]
        print("%-86s %7d  %s" % (n[:86], c["_total"], "  ".join("%*d" % (len(p), c[p]) for p in cols)))
        tot.update(c)
    print("%-86s %7d  %s" % ("TOTAL", tot["_total"], "  ".join("%*d" % (len(p), tot[p]) for p in PATS if p not in ("STG", "LDG", "LDS", "STS", "DADD", "DMUL"))))
    print("# HMMA / IMMA (legacy mma.sync) instructions in the library: %d (every tensor-core kernel is tcgen05 + TMEM + TMA)" % (tot["HMMA"] + tot["IMMA"]))


if __name__ == "__main__":
    sys.exit(main())
