"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel time of one simulation step.

A step starts at k_bp_separate.  If the capture window holds two such launches the launches between them are the
step; if it holds one (the window straddles a step boundary) the step is stitched from the head after it and the
matching tail of the previous step."""
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
note = ""
if len(idx) >= 2:
    step = seq[idx[0]:idx[1]]
    # the e2e upload path (pack / unpack / refresh) sits between two steps of bench.py's e2e loop; keep only the step itself
    while step and any(k in step[-1][0] for k in ('k_pack_state', 'k_unpack_state', 'k_refresh_bodies')):
        step.pop()
else:
    head = seq[idx[0]:]
    tail = [x for x in seq[:idx[0]] if not any(k in x[0] for k in ('k_pack_state', 'k_unpack_state', 'k_refresh_bodies'))]
    names, key = [n for n, _ in tail], [n for n, _ in head[-3:]]
    pos = [i for i in range(len(tail) - 2) if names[i:i + 3] == key]
    step = head + tail[pos[0] + 3:]
    note = "; stitched from two consecutive steps of one capture"
tot = sum(v for _, v in step)
agg = collections.OrderedDict()
for n, v in step:
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v
print(f"one step = {len(step)} launches, {tot:.1f} us (ncu per-launch times: cold cache, serialised{note})")
for n, (c, v) in agg.items():
    print(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{c:<2d} {n}")
