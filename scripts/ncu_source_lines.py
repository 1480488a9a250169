"""Per CUDA source line of one kernel: warp instructions executed, stall samples (total / long scoreboard / no instruction),
from an .ncu-rep captured with --import-source on.  usage: python scripts/ncu_source_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"; hdr = None; cur = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = {}; [hdr.setdefault(h, i) for i, h in enumerate(r)]; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "": cur = (fname, int(r[0]), r[1].strip()[:110]); agg.setdefault(cur, [0, 0, 0, 0, 0]); continue
    if cur is None or r[2] == "...": continue
    def f(k):
        try: return float(r[hdr[k]] or 0)
        except (ValueError, KeyError): return 0.0
    a = agg[cur]; a[0] += f("Instructions Executed"); a[1] += f("# Samples"); a[2] += f("stall_long_sb"); a[3] += f("stall_no_inst"); a[4] += 1
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("warp instructions %.4g, samples %d, static SASS %d" % (ti, ts, sum(a[4] for a in agg.values())))
print("%-22s %6s %6s %6s %6s %5s  source" % ("file:line", "inst%", "smpl%", "lsb%", "noin%", "sass"))
for (fn, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print("%-22s %6.2f %6.2f %6.2f %6.2f %5d  %s" % ("%s:%d" % (fn, ln), 100 * a[0] / max(ti, 1), 100 * a[1] / max(ts, 1), 100 * a[2] / max(ts, 1), 100 * a[3] / max(ts, 1), a[4], src))
