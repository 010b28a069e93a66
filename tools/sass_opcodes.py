#!/usr/bin/env python
"""Opcode histogram of the built library, per kernel: the SASS evidence for the tcgen05 / TMEM / TMA claims.

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt

Counts the Blackwell-specific mnemonics (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk,
UTMALDG / UTMASTG = tensor-map TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and the arithmetic that dominates the
search epilogue (FMNMX3 / VIMNMX3), plus HMMA (legacy mma.sync; must be absent)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fractal-image-compression_b200", "lib", "libfic_b200.so")
WATCH = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FMNMX3", "VIMNMX3",
         "HMMA", "IMMA", "IDP4A", "IDP", "ATOMG", "RED", "LDG", "STG", "LDS", "STS"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    names = demangle(list(kernels))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a; {len(kernels)} kernels)")
    print("# kernel | instructions | " + " ".join(WATCH))
    for k, c in kernels.items():
        short = re.sub(r"fic::\(anonymous namespace\)::|fic::", "", names[k])
        short = re.sub(r"\(.*", "", short)
        cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{short:60s} {c['_total']:6d}  {cols}")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# library total: " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
    if tot["HMMA"] or tot["IMMA"]:
        print("# WARNING: legacy mma.sync instructions present")


if __name__ == "__main__":
    sys.exit(main())
