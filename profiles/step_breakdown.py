"""Summarise one training step from an ncu launch list (gpu__time_duration.sum CSV):
    python profiles/step_breakdown.py launches.csv [step_index]"""
import csv, re, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
data = [(x[ki], float(x[vi].replace(',', ''))) for x in r]
# a step starts with the flat-gradient zero fill that precedes the first Chebyshev launch of layer 1
idx = [i for i, (n, _) in enumerate(data) if 'cheb_fwd' in n]
per_step = 16
starts = idx[::per_step]
step = data[starts[which]:starts[which + 1]]
agg = {}
for n, v in step:
    n = re.sub(r'^void ', '', re.sub(r'\(.*', '', n))[:72]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in step)
print('launches in step', len(step), 'serialised us %.1f' % (tot / 1e3))
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print('%7.1f us %5.1f%% x%3d %s' % (v / 1e3, 100 * v / tot, c, n))
