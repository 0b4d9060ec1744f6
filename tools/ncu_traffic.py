#!/usr/bin/env python
"""profiles/gemm_traffic.json from an ncu launch list of one bench step (run here, no GPU needed):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 ...   (on the GPU box)
    python tools/ncu_traffic.py gpurun_out/launches.csv profiles/r02_launches.csv

Writes the mean DRAM bytes per GEMM launch of the LAST full step in the list together with the SHA-256 of the GEMM
sources the library was built from; bench.py prints `traffic: null` when that hash is not the current tree's."""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import GEMM_SOURCES, source_sha256  # noqa: E402


def main():
    src, keep = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iID = hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(int(r[iID]), {"name": r[iK]})
        d[r[iM]] = float(r[iV].replace(",", ""))
    launches = list(per.values())
    gemm = [d for d in launches if "gemm_bf16" in d["name"]]
    step = gemm[-121:] if len(gemm) >= 121 else gemm  # 115 tower + 4 head + 2 projector launches per step
    n = len(step)
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in step) / n
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in step) / n
    fam = collections.Counter()
    for d in launches:
        fam[d["name"].split("(")[0].split("<")[0]] += d.get("gpu__time_duration.sum", 0.0)
    total = sum(fam.values()) or 1.0
    out = {"dram_bytes_per_launch": round(rd + wr), "dram_read_bytes_per_launch": round(rd),
           "dram_write_bytes_per_launch": round(wr), "launches": n, "source_sha256": source_sha256(*GEMM_SOURCES),
           "capture": os.path.basename(keep or src),
           "time_share_by_kernel": {k: round(v / total, 4) for k, v in fam.most_common(12)},
           "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                  "over bench.py steps; mean over the GEMM launches of the last step in the list"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "gemm_traffic.json"), "w"), indent=1)
    if keep:
        shutil.copyfile(src, keep)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
