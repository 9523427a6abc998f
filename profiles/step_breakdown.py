"""Summarise one training step from an ncu launch list (gpu__time_duration.sum CSV):
    python profiles/step_breakdown.py launches.csv [step_index]
A step is delimited by bench.py's L2 flush (the 256 MB fill that precedes every timed step)."""
import csv, re, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
data = [(x[ki], float(x[vi].replace(',', ''))) for x in r]
flush = [i for i, (n, v) in enumerate(data) if 'FillFunctor' in n and v > 25000]
print('flush fills at', flush)
a = flush[which - 1] if which != -1 else flush[-2]
b = flush[which] if which != -1 else flush[-1]
step = data[a + 1:b]
agg = {}
for n, v in step:
    n = re.sub(r'^void ', '', re.sub(r'\(.*', '', n))[:72]
    e = agg.setdefault(n, [0, 0.0]); e[0] += 1; e[1] += v
tot = sum(v for _, v in step)
ours = sum(v for n, v in step if 'agcn::' in n)
print('launches in step', len(step), 'serialised us %.1f' % (tot / 1e3), 'agcn kernels %.1f%%' % (100 * ours / tot))
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print('%7.1f us %5.1f%% x%3d %s' % (v / 1e3, 100 * v / tot, c, n))
