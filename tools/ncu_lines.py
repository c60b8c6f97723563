#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export per source line / per function range."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
tot = collections.Counter(); samp = collections.Counter(); text = {}
cur = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]
    elif r and r[0] == "Line No": hdr = r
    elif r and r[0].isdigit() and cur:
        d = dict(zip(hdr, r))
        try: n = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
        except Exception: n = s = 0
        key = (cur, int(r[0])); tot[key] += n; samp[key] += s; text[key] = r[1]
T = sum(tot.values()); S = sum(samp.values())
print("total inst", T, "samples", S)
mode = sys.argv[2] if len(sys.argv) > 2 else "inst"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
src = tot if mode == "inst" else samp
for k, v in src.most_common(top):
    print(f"{k[0][5:]}:{k[1]:4d} inst {100*tot[k]/T:5.2f}% samp {100*samp[k]/S:5.2f}%  {text[k].strip()[:105]}")
