"""SASS (with execution counts) attributed to a source-line range.  usage: sass_of_lines.py file.csv first last [file]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); lo, hi = int(sys.argv[2]), int(sys.argv[3]); want = sys.argv[4] if len(sys.argv) > 4 else 'st.cu'
h = None; cur = None; fname = None; seen = set(); out = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): fname = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": h = r; continue
    if h is None or len(r) != len(h): continue
    if r[0].isdigit(): cur = int(r[0]); continue
    if fname == want and cur is not None and lo <= cur <= hi and r[2] not in seen:
        seen.add(r[2])
        try: n = int(r[7])
        except ValueError: n = -1
        out.append((r[2], cur, r[3].strip(), n))
for a, l, s, n in sorted(out): print(a[-5:], l, "%9d" % n, s)
