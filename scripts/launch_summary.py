"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals, shares
and the per-round progression of the staged solver.  usage: launch_summary.py launches.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; data = rows[hi + 1:]
ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
scale = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}
agg = collections.OrderedDict(); seq = []
for r in data:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
    name = r[ki].split('(')[0].replace('mmpc::', '')
    agg.setdefault(name, []).append(v); seq.append((name, v))
tot = sum(sum(v) for v in agg.values())
print("total %.2f ms over %d launches" % (tot, len(seq)))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-34s n=%5d sum=%9.2f ms share=%5.1f%% first=%8.3f median=%8.3f min=%8.3f" % (
        k[:34], len(v), sum(v), 100 * sum(v) / tot, v[0], sorted(v)[len(v) // 2], min(v)))
names = [n for n, _ in seq]
if 'staged_eval_kernel' in names:
    per = {}
    for n, v in seq:
        if n.startswith('staged_') and n != 'staged_init_kernel': per.setdefault(n, []).append(v)
    nr = min(len(v) for v in per.values() if len(v) > 2)
    keys = [k for k in per if len(per[k]) >= nr]
    print("round " + " ".join(k.replace('staged_', '').replace('_kernel', '')[:10].rjust(10) for k in keys))
    for r in (0, 1, 2, 5, 10, 20, 30, 40, 50, 60, 80, 100, 150, 200, 250):
        if r < nr: print("%5d " % r + " ".join(("%.3f" % per[k][r * (len(per[k]) // nr)]).rjust(10) for k in keys))
