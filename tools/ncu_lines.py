"""Per-source-line instruction and stall-sample shares of one kernel of an .ncu-rep (captured with --import-source on).
Usage: python tools/ncu_lines.py report.ncu-rep <kernel-name-substring> [launch-index-among-matches] [top N]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# split into (function, file) blocks; a kernel launch = consecutive blocks with the same function name
blocks, cur = [], None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = {"file": r[1], "fn": None, "hdr": None, "rows": []}; blocks.append(cur); continue
    if r[0] == "Function Name": cur["fn"] = r[1]; continue
    if r[0] == "Line No": cur["hdr"] = r; continue
    if cur is not None and cur["hdr"] is not None: cur["rows"].append(r)
launches, last = [], None
for b in blocks:
    if b["fn"] != last or (launches and any(x["file"] == b["file"] for x in launches[-1])):
        launches.append([]); last = b["fn"]
    launches[-1].append(b)
sel = [l for l in launches if pat in l[0]["fn"]]
L = sel[which]
print("kernel:", L[0]["fn"][:100], "(%d matching launches)" % len(sel))
agg = collections.OrderedDict()
tot_i = tot_s = 0
for b in L:
    h = b["hdr"]
    ii, si, ti = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    for r in b["rows"]:
        if r[0] == "": continue   # SASS rows
        key = (b["file"].split("/")[-1], int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0, 0])
        n = len(h)   # a source line with quotes / commas is split into extra fields: index the counters from the end
        a[0] += int(r[ii - n]); a[1] += int(r[si - n]); a[2] += int(r[ti - n])
for a in agg.values(): tot_i += a[0]; tot_s += a[1]
print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
print("%-14s %5s %7s %7s %6s  %s" % ("file", "line", "inst%", "samp%", "thr/i", "source"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%-14s %5d %6.2f%% %6.2f%% %6.1f  %s" % (k[0], k[1], 100.0 * a[0] / max(tot_i, 1), 100.0 * a[1] / max(tot_s, 1), a[2] / max(a[0], 1), k[2]))
# per file summary
pf = collections.defaultdict(lambda: [0, 0])
for k, a in agg.items(): pf[k[0]][0] += a[0]; pf[k[0]][1] += a[1]
for f, a in sorted(pf.items(), key=lambda x: -x[1][1]): print("FILE %-16s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * a[0] / max(tot_i, 1), 100.0 * a[1] / max(tot_s, 1)))
# optional: coarse groups by (file, line range) given as extra args "name=file:lo-hi"
groups = [a for a in sys.argv[5:] if "=" in a]
if groups:
    print("--- groups")
    for gspec in groups:
        name, rest = gspec.split("=")
        f, rng = rest.split(":")
        lo, hi = map(int, rng.split("-"))
        gi = sum(a[0] for k, a in agg.items() if k[0] == f and lo <= k[1] <= hi)
        gs = sum(a[1] for k, a in agg.items() if k[0] == f and lo <= k[1] <= hi)
        print("%-28s inst %5.1f%%  samples %5.1f%%" % (name, 100.0 * gi / max(tot_i, 1), 100.0 * gs / max(tot_s, 1)))
