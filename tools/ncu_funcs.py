#!/usr/bin/env python
"""Per-function dynamic instruction totals from an `ncu --page source --csv` export.
usage: ncu_funcs.py src.csv <git-rev of the profiled sources> [steps_per_launch]"""
import csv, collections, os, re, subprocess, sys
rows = list(csv.reader(open(sys.argv[1])))
rev = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
units = float(sys.argv[3]) if len(sys.argv) > 3 else 4096 * 17
tot = collections.Counter(); samp = collections.Counter()
cur = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1]
    elif r and r[0] == "Line No": hdr = r
    elif r and r[0].isdigit() and cur:
        d = dict(zip(hdr, r))
        try: n = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
        except Exception: n = s = 0
        tot[(cur, int(r[0]))] += n; samp[(cur, int(r[0]))] += s
fn_of = {}
for path in {k[0] for k in tot}:
    rel = path.split("/root/repo/")[-1]
    try: text = subprocess.run(["git", "show", f"{rev}:{rel}"], capture_output=True, text=True, cwd="/root/repo").stdout.split("\n")
    except Exception: text = []
    if len(text) < 5:  # no git history next to the sources (the GPU box): the working tree is what was profiled
        try: text = open(os.path.join("/root/repo", rel)).read().split("\n")
        except Exception: text = []
    name = "?"
    starts = []
    for i, line in enumerate(text, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:RCSB_DEV(?:_NOINLINE)?|__global__|__device__|static|inline)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line)
        if m and not line.strip().startswith("//"): starts.append((i, m.group(1)))
    fn_of[path] = starts
agg = collections.Counter(); agg_s = collections.Counter()
for (path, ln), n in tot.items():
    name = "?"
    for s, nm in fn_of.get(path, []):
        if s <= ln: name = nm
        else: break
    key = path.split("/")[-1].replace("rcsb_", "") + ":" + name
    agg[key] += n; agg_s[key] += samp[(path, ln)]
T = sum(agg.values()); S = sum(agg_s.values())
print(f"total {T} instructions = {T/units:.0f} per env physics step; {S} samples")
for k, v in agg.most_common(45):
    print(f"{k:45s} {v/units:8.0f} inst/step {100*v/T:5.1f}%   samples {100*agg_s[k]/S:5.1f}%")
