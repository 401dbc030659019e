"""Summarise an ncu launch list (gpu__time_duration.sum) for one d_sw call: per-kernel time and share."""
import csv, sys
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith('==')]))
names = [(x['Kernel Name'], float(x['Metric Value'])) for x in rows if x['Metric Name'] == 'gpu__time_duration.sum']
idx = [i for i, (n, _) in enumerate(names) if 'k_dsw_wind' in n]
seq = names[idx[-2]:idx[-1]] if len(idx) >= 2 else names[idx[-1]:]
tot = sum(v for _, v in seq)
print(f"d_sw: {len(seq)} launches, {tot/1e6:.3f} ms (ncu, cold-cache serialised)")
agg = {}
for n, v in seq:
    k = n.split('(')[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (cnt, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:16s} x{cnt:2d} {v/1e3:9.1f} us {100*v/tot:5.1f}%")
