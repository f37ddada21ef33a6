"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown share table.
usage: python profiles/summarize_launches.py gpurun_out/<file>.csv "<title>" "<command>" > profiles/<name>.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[h]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
d, tot = collections.OrderedDict(), 0.0
for r in rows[h + 1:]:
    if len(r) > mv:
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "")[:90]
        t = float(r[mv].replace(",", ""))
        d.setdefault(name, []).append(t)
        tot += t
print(f"# {sys.argv[2]}\n\nCommand: `{sys.argv[3]}`\n")
print(f"{sum(len(v) for v in d.values())} launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare SHARES).\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v) / 1e3:.1f} |")
