"""Side-by-side table of tests/diag_gemm_shapes.py outputs.  usage: python profiles/compare_shapes.py name=file ..."""
import sys


def load(f):
    d = {}
    for l in open(f):
        p = l.split()
        if len(p) >= 8 and p[-1].replace('.', '').isdigit() and not l.startswith('total') and not l.startswith('gemm'):
            d[' '.join(p[:-7])] = (float(p[-4]), int(p[-2]))
    return d


fs = dict(a.split('=', 1) for a in sys.argv[1:])
D = {k: load(v) for k, v in fs.items()}
first = next(iter(D.values()))
print('%-18s' % 'gemm (us/launch)', *['%11s' % k for k in fs], '  per step')
tot = {k: 0 for k in fs}
best = 0
for n in first:
    row = [D[k].get(n, (float('nan'), 0))[0] for k in fs]
    mult = first[n][1]
    for k, v in zip(fs, row):
        tot[k] += v * mult
    best += min(row) * mult
    print('%-18s' % n, *['%11.1f' % v for v in row], '  x%d' % mult)
print('%-18s' % 'total us/step', *['%11.1f' % tot[k] for k in fs])
print('best kernel per shape: %.1f us/step' % best)
