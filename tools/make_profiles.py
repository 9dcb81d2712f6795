"""Turns the files a `tools/gpu_round2.sh <tag>` visit left in gpurun_out/ into the tracked summaries under profiles/:
bench lines, the ncu launch list + per-kernel shares, k_trace DRAM / L2 traffic per launch (profiles/traffic.json, read by
bench.py for roofline.traffic), and the headline metrics of the --set full captures.
Usage: python tools/make_profiles.py <tag>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def rows_of(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1:]
    raise SystemExit("no header in " + path)


for a, b in (("bench_%s.json", "%s_bench.json"), ("bench_ref_%s.json", "%s_bench_reference.json"), ("launches_%s.csv", "%s_launches_bench.csv"),
             ("trace_traffic_%s.csv", "%s_trace_traffic.csv")):
    shutil.copy(os.path.join(G, a % tag), os.path.join(P, b % tag))

h, rows = rows_of(os.path.join(G, "launches_%s.csv" % tag))
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
    n = r[ki].split("(")[0][:60]
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
lines = ["%-60s %5d launches %10.3f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1][1])]
open(os.path.join(P, "%s_launches_bench_summary.txt" % tag), "w").write(
    "ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (whole process: counting pass, "
    "3 warm-up steps, 1 timed step, 3 end-to-end steps; cold-cache serialized per-launch times)\n" + "\n".join(lines) + "\n")
print("\n".join(lines[:8]))

h, rows = rows_of(os.path.join(G, "trace_traffic_%s.csv" % tag))
mi, vi, ui, idi = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
per = collections.defaultdict(dict)
for r in rows:
    if len(r) <= vi:
        continue
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(r[ui], 1)
    per[r[idi]][r[mi]] = float(r[vi].replace(",", "")) * mult
n = len(per)
rd = sum(p["dram__bytes_read.sum"] for p in per.values())
wr = sum(p["dram__bytes_write.sum"] for p in per.values())
l2 = sum(p["lts__t_bytes.sum"] for p in per.values())
b = json.load(open(os.path.join(G, "bench_%s.json" % tag)))
alg = b["roofline"]["bytes_per_ray"] * b["value"] * 1e6 * b["ms_per_step"] / 1e3 / b["roofline"]["trace_launches_per_step"]
# profiles/traffic.json is keyed by bench config ("config5" = BASELINE configs[4], ...): bench.py reports the entry of the config it runs
import re
cfg_key = "config%d" % (int(re.search(r"configs\[(\d)\]", b["config"]["workload"]).group(1)) + 1)
tpath = os.path.join(P, "traffic.json")
allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
if "k_trace_dram_bytes_per_launch" in allt:      # the round-1 flat file was measured on config 2
    allt = {"config2": allt}
allt[cfg_key] = {"k_trace_dram_bytes_per_launch": (rd + wr) / n, "k_trace_dram_read_bytes_per_launch": rd / n, "k_trace_dram_write_bytes_per_launch": wr / n,
                 "k_trace_l2_bytes_per_launch": l2 / n, "k_trace_algorithmic_bytes_per_launch": alg, "launches": n,
                 "source": "profiles/%s_trace_traffic.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum over the %d k_trace launches of the "
                           "timed step of `bench.py --steps 1 --warmup 3 --no-cpu-baseline` (tools/gpu_r2final.sh)" % (tag, n)}
json.dump(allt, open(tpath, "w"), indent=1)
print("k_trace per launch: DRAM %.0f MB, L2 %.0f MB, algorithmic %.0f MB (%d launches)" % ((rd + wr) / n / 1e6, l2 / n / 1e6, alg / 1e6, n))
for kind, width in (("trace", 260), ("shade", 420)):
    if not os.path.exists(os.path.join(G, "prof_%s_%s.ncu-rep" % (kind, tag))):
        print("no prof_%s_%s.ncu-rep: summary skipped" % (kind, tag))
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "prof_%s_%s.ncu-rep" % (kind, tag))], capture_output=True, text=True).stdout
    open(os.path.join(P, "%s_ncu_%s_summary.txt" % (tag, kind)), "w").write("\n".join(l[:width] for l in out.splitlines()) + "\n")
