"""Per-source-line instruction share from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: python profiles/source_breakdown.py file.csv [min_pct]
Inlined helpers are counted at their definition AND at the call site, so percentages are per file, not additive across files."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
per = collections.OrderedDict(); src = {}; stall = collections.Counter()
fname = None; h = None
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": h = r; continue
    if h is None or len(r) != len(h): continue
    if not r[0].isdigit(): continue   # SASS rows nested under their source line
    d = dict(zip(h, r))
    try: n = int(d["Instructions Executed"])
    except ValueError: continue
    key = (fname, int(r[0]))
    per[key] = per.get(key, 0) + n
    src[key] = r[1]
    try: stall[key] += int(d["# Samples"])
    except ValueError: pass
main = max(set(k[0] for k in per), key=lambda f: sum(n for k, n in per.items() if k[0] == f))
per = collections.OrderedDict((k, n) for k, n in per.items() if k[0] == main)
tot = sum(per.values()); ts = sum(stall[k] for k in per) or 1
print("total warp instructions", tot)
for k, n in per.items():
    if 100.0 * n / tot >= minp or 100.0 * stall[k] / ts >= minp:
        print("%-14s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0][:14], k[1], 100.0 * n / tot, 100.0 * stall[k] / ts, src[k].strip()[:100]))
