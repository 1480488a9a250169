"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line and per
solver phase.  usage: python scripts/ncu_lines.py dump.csv [top_n]"""
import csv, re, sys, os
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur = None; hdr = None; recs = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0] != '':
        d = dict(zip(hdr[4:], r[4:])); recs.append((cur, int(r[0]), r[1], d))
def fl(d, k):
    try: return float(d.get(k, '0') or 0)
    except ValueError: return 0.0
tot_i = sum(fl(d, 'Instructions Executed') for _, _, _, d in recs)
tot_s = sum(fl(d, '# Samples') for _, _, _, d in recs)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, 'mobile_manipulator_mpc_b200/csrc/mmpc_lane.cuh')).read().split('\n')
marks = [(i + 1, re.sub(r'\(.*', '', l.strip().replace('__device__ ', '').replace('__forceinline__ ', '')))
         for i, l in enumerate(src) if re.match(r'\s*__device__ .*\(', l) and not l.strip().startswith('//')]
def phase(f, line):
    if not f.endswith('mmpc_lane.cuh'): return os.path.basename(f)
    name = '?'
    for ln, l in marks:
        if ln <= line: name = l
    return name
agg = {}
for f, ln, s, d in recs:
    k = phase(f, ln); a = agg.setdefault(k, [0, 0, {}])
    a[0] += fl(d, 'Instructions Executed'); a[1] += fl(d, '# Samples')
    for st in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_lg', 'stall_mio', 'stall_selected', 'stall_not_selected', 'stall_no_inst', 'stall_branch_resolving', 'stall_dispatch'):
        a[2][st] = a[2].get(st, 0) + fl(d, st)
print("total warp instructions %.4g, samples %d" % (tot_i, tot_s))
print("== per phase: %inst %samples  top stalls")
for k, (i, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print("%5.1f%% %5.1f%%  %-44s %s" % (100 * i / tot_i, 100 * s / max(tot_s, 1), k[:44], ' '.join('%s=%.0f%%' % (a.replace('stall_', ''), 100 * b / max(s, 1)) for a, b in top)))
print("== top lines by samples")
for f, ln, s, d in sorted(recs, key=lambda r: -fl(r[3], '# Samples'))[:topn]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * fl(d, '# Samples') / max(tot_s, 1), 100 * fl(d, 'Instructions Executed') / tot_i, os.path.basename(f), ln, s.strip()[:90]))
