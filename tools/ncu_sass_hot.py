#!/usr/bin/env python
"""Per-SASS-instruction executed counts of one kernel in an ncu report, grouped into contiguous regions.
    python tools/ncu_sass_hot.py rep.ncu-rep <kernel substring> [min_share]"""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, kernels = None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; kernels[cur] = []
    elif r and r[0].startswith("0x") and cur is not None:
        kernels[cur].append(r)
hdr = next(r for r in rows if r and r[0] == "Address")
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
for name, ins in kernels.items():
    if pat not in name:
        continue
    tot = sum(int(r[ci]) for r in ins); tots = sum(int(r[cs]) for r in ins)
    print(name, "total warp instr", tot, "samples", tots)
    for i, r in enumerate(ins):
        n = int(r[ci])
        print(f"{i:5d} {n/tot*100:6.2f}% smp {int(r[cs])/max(tots,1)*100:5.2f}% lanes {int(r[ct])/max(n,1):5.1f}  {r[1].strip()[:90]}")
    break
