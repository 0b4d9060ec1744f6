#!/usr/bin/env python
"""Opcode evidence per kernel of the in-tree library (run here, no GPU needed): counts of the SASS mnemonics that
prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA ->
UTMALDG/UTMASTG, legacy mma.sync -> HMMA) plus registers / shared memory from the ELF resource usage.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gameplay_vision_llm_b200", "libgvl_sm100a.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS",
         "HMMA", "IDP", "MUFU.EX2", "DFMA", "FFMA2", "ACQBULK", "UCGABAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), sm_100a")
    print("# tcgen05.mma -> UTCHMMA(.2CTA = cta_group::2), tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG; HMMA = legacy mma.sync")
    tot = collections.Counter()
    for (mangled, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name)
        regs, smem = usage.get(mangled, (0, 0))
        hits = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{short[:110]:110s} insts={c['_total']:6d} regs={regs:3d} static_smem={smem:6d}  {hits}")
        tot.update({w: c[w] for w in WATCH})
    print("# library totals: " + " ".join(f"{w}={tot[w]}" for w in WATCH))


if __name__ == "__main__":
    main()
