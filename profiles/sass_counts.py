"""Static instruction-class counts of every kernel in libsurfel_b200.so (cuobjdump -sass), so that the claims
about the code path (LDGSTS / cp.async staging, no TMA on the production path, one REDG per column lane, no
tensor-core instruction anywhere, no local-memory spills in the hot kernels) can be checked without
rebuilding.   python profiles/sass_counts.py [lib.so] > profiles/r2_sass_counts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "2d-gaussian-splatting_b200", "lib", "libsurfel_b200.so")
CLASSES = [("LDG", r"^LDG"), ("LDGSTS (cp.async)", r"^LDGSTS"), ("UBLKCP/UTMA (TMA)", r"^(UBLKCP|UTMA)"),
           ("LDS", r"^LDS"), ("STS", r"^STS"), ("STG", r"^STG"), ("REDG", r"^RED"), ("ATOMG/ATOMS", r"^ATOM"),
           ("LDL/STL (spill)", r"^(LDL|STL)"), ("FFMA/FMUL/FADD", r"^(FFMA|FMUL|FADD)\b"), ("FFMA2/FMUL2/FADD2", r"^(FFMA2|FMUL2|FADD2)"),
           ("DFMA/DMUL/DADD", r"^(DFMA|DMUL|DADD)"), ("MUFU", r"^MUFU"), ("FSETP/FMNMX/FSEL", r"^(FSETP|FMNMX|FSEL)"),
           ("VOTE", r"^VOTE"), ("SHFL", r"^SHFL"), ("FLO/BMSK/POPC", r"^(FLO|BMSK|POPC)"), ("BAR", r"^BAR"),
           ("BRA/BRX/BSSY/BSYNC", r"^(BRA|BRX|BSSY|BSYNC)"), ("HMMA/UTC*MMA (tensor)", r"^(HMMA|UTC|IMMA|QMMA)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern).replace("surfel::", "")
            counts[kern] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and kern:
            op = m.group(1)
            total[kern] += 1
            for name, pat in CLASSES:
                if re.match(pat, op):
                    counts[kern][name] += 1
    names = [n for n, _ in CLASSES]
    print("# SASS instruction-class counts per kernel (static; `cuobjdump -sass lib/libsurfel_b200.so`)\n")
    print("| kernel | total | " + " | ".join(names) + " |")
    print("|---|---|" + "---|" * len(names))
    for k, c in counts.items():
        print(f"| `{k}` | {total[k]} | " + " | ".join(str(c.get(n, 0)) for n in names) + " |")


if __name__ == "__main__":
    main()
