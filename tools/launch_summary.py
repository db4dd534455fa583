"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel time of one simulation step."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
seq = []
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    seq.append((r[ki].split('(')[0][:70], v))
idx = [i for i, (n, _) in enumerate(seq) if 'k_bp_separate' in n]
a, b = idx[0], idx[1]
tot = sum(v for _, v in seq[a:b])
agg = collections.OrderedDict()
for n, v in seq[a:b]:
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v
print(f"one step = {b - a} launches, {tot:.1f} us (ncu per-launch times: cold cache, serialised)")
for n, (c, v) in agg.items():
    print(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{c:<2d} {n}")
