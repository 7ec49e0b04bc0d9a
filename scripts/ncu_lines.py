#!/usr/bin/env python
"""Per-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: scripts/ncu_lines.py rep.ncu-rep [top N]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; H = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
for r in rows:
  if not r: continue
  if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
  if r[0] == "Line No": H = r; iS = H.index("# Samples"); iI = H.index("Instructions Executed"); st = [i for i, x in enumerate(H) if x.startswith("stall_") and "Not Issued" not in x]; continue
  if H is None or len(r) != len(H) or not r[0].isdigit(): continue
  key = (cur, int(r[0]))
  a = agg[key]
  if r[1].strip(): a[3] = r[1].strip()
  if r[iS].isdigit(): a[0] += int(r[iS])
  if r[iI].isdigit(): a[1] += int(r[iI])
  for i in st:
    if r[i] not in ("", "0"): a[2][H[i]] += int(r[i])
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]
for (f, ln), a in sorted(top):
  print(f"{f}:{ln:<5d} {a[0]:6d} {100*a[0]/max(tot,1):5.1f}% inst={a[1]/1e6:7.2f}M {dict(a[2].most_common(2))} | {a[3][:90]}")
