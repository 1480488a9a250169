"""Top SASS instructions of one kernel by a stall reason, from an .ncu-rep with source counters.
usage: python scripts/ncu_sass_stalls.py report.ncu-rep kernel_regex [stall_long_sb] [top_n]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
col = sys.argv[3] if len(sys.argv) > 3 else "stall_long_sb"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; recs = []; seen = set()
for r in rows:
    if r and r[0] == "Address": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] in seen: continue
    seen.add(r[0])
    def f(k):
        try: return float(r[hdr[k]] or 0)
        except (ValueError, KeyError): return 0.0
    recs.append((len(recs), r[1].strip(), f("# Samples"), f(col), f("Instructions Executed")))
ts = sum(x[2] for x in recs); tl = sum(x[3] for x in recs); ti = sum(x[4] for x in recs)
print("%d SASS instructions, %d samples, %s %d (%.1f%%), warp instructions executed %.4g" % (len(recs), ts, col, tl, 100 * tl / max(ts, 1), ti))
op = collections.Counter()
for x in recs:
    w = x[1].split(); op[w[1] if w and w[0].startswith("@") else (w[0] if w else "?")] += x[3]
print("by opcode of the stalled instruction:", ", ".join("%s %.1f%%" % (k, 100 * v / max(tl, 1)) for k, v in op.most_common(10)))
for x in sorted(recs, key=lambda x: -x[3])[:topn]:
    print("%6d %5.2f%%  exec=%-8d %s" % (x[0], 100 * x[3] / max(tl, 1), x[4], x[1][:100]))
