#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): headline raw metrics, opcode mix and hot SASS regions.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
       "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
       "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    d = data[which]
    print("kernel:", d[hdr.index("Kernel Name")][:100])
    for m in RAW:
        if m in hdr:
            i = hdr.index(m)
            print(f"  {m:75s} {d[i]:>16s} {units[i]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                v = float(d[i])
            except ValueError:
                continue
            if v >= 0.15:
                print(f"  stall {h[34:-23]:40s} {v:6.2f}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    # the source page concatenates kernels; pick the which-th block
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    blk = blocks[min(which, len(blocks) - 1)]
    h = blk[0]
    iS, iI, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ops, smp = collections.Counter(), collections.Counter()
    seg, cur = [], None
    tot = 0
    for r in blk[1:]:
        if len(r) < 10:
            continue
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "IDP", "LDTM", "STTM")) else op.split(".")[0]
        n, sm = int(r[iI]), int(r[iSm] or 0)
        ops[op] += n
        smp[op] += sm
        tot += n
        if cur is None or abs(n - cur[2]) > 0.02 * max(n, cur[2], 1):
            if cur:
                seg.append(cur)
            cur = [r[0][-5:], r[0][-5:], n, 0, 0, 0]
        cur[1] = r[0][-5:]
        cur[3] += 1
        cur[4] += n
        cur[5] += sm
    seg.append(cur)
    ts = max(1, sum(s[5] for s in seg))
    print(f"  total warp instructions {tot}")
    for o, n in ops.most_common(16):
        print(f"    {o:14s} {n:>12d} {n / tot * 100:5.1f}%  samples {smp[o] / ts * 100:5.1f}%")
    print("  hot regions (address range, executions per instruction, instructions, share of instr / samples)")
    for s in seg:
        if s[4] > tot * 0.01 or s[5] > ts * 0.02:
            print(f"    {s[0]}-{s[1]} x{s[2]:>9d} n={s[3]:>4d} {s[4] / tot * 100:5.1f}% / {s[5] / ts * 100:5.1f}%")


if __name__ == "__main__":
    main()
