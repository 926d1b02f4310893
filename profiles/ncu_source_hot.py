#!/usr/bin/env python
"""Top CUDA source lines by warp-stall samples from an .ncu-rep captured with --import-source on.
usage: python profiles/ncu_source_hot.py rep.ncu-rep [n_lines]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; fname = ""; agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not r[0]: continue
    agg.append((fname, r))
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
names = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for _, r in agg); toti = sum(int(r[ie] or 0) for _, r in agg)
print(f"total samples {tot}, warp instructions {toti}")
for f, r in sorted(agg, key=lambda t: -int(t[1][si] or 0))[:topn]:
    st = sorted(((int(r[i] or 0), n) for i, n in names), reverse=True)[:3]
    print("%5.1f%% smp %5.1f%% inst  %s:%s | %-90s | %s" % (100 * int(r[si]) / tot, 100 * int(r[ie] or 0) / toti, f, r[0],
          r[1].strip()[:90], " ".join(f"{n}={v}" for v, n in st)))
