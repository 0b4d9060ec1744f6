#!/usr/bin/env python
"""Top stall sites of an .ncu-rep (source page, SASS): python tools/ncu_hot.py rep [kernel-index] [top-n]
Prints the instructions with the most warp-stall samples together with their neighbours' opcodes."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    blk = blocks[min(which, len(blocks) - 1)]
    h = blk[0]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    rows = [r for r in blk[1:] if len(r) > iE]
    tot = sum(int(r[iN] or 0) for r in rows) or 1
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][iN] or 0))[:top]
    print(f"total samples {tot}, instructions {len(rows)}")
    for i in sorted(order):
        r = rows[i]
        prev = rows[i - 1][iS].split()[:2] if i else []
        print(f"  #{i:5d} {int(r[iN] or 0) / tot * 100:5.1f}%  x{int(r[iE]):>9d}  {r[iS].strip()[:70]:70s}  <- {' '.join(prev)[:40]}")


if __name__ == "__main__":
    main()
