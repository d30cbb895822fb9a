"""Summarise an `ncu --csv --log-file` launch list (metrics gpu__time_duration.sum [+ sm__pipe_tensor_cycles_active...
pct_of_peak_sustained_elapsed]): per-kernel launch counts, summed time, share; time-weighted tensor-pipe activity of
the conv_tc launches.  usage: summarize_ncu.py launches.csv [first_launch_index]   (index: skip the warm-up forward)"""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
per_id = collections.OrderedDict()
for r in rows:
    d = per_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    d[r["Metric Name"]] = v
launches = list(per_id.values())
start = int(sys.argv[2]) if len(sys.argv) > 2 else len(launches) // 2
launches = launches[start:]
tot = sum(l.get("gpu__time_duration.sum", 0.0) for l in launches)
print("%d launches (from index %d), %.2f ms summed (cold-cache serialised ncu times: compare shares)" % (len(launches), start, tot))
agg = collections.OrderedDict()
for l in launches:
    name = re.sub(r"\(.*", "", l["name"]).replace("void ", "").replace("pcnn::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    t = l.get("gpu__time_duration.sum", 0.0)
    a[0] += 1; a[1] += t
    a[2] += t * l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
for name, (n, t, tw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = "   tensor pipe active %.1f %%" % (tw / t) if tw > 0 and t > 0 else ""
    print("%-60s n=%4d %9.3f ms %5.1f%%%s" % (name[:60], n, t, 100 * t / tot, extra))
conv = [l for l in launches if "conv_tc_kernel" in l["name"]]
key = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
if conv and any(key in l for l in conv):
    ct = sum(l["gpu__time_duration.sum"] for l in conv)
    print("conv_tc: %d launches, %.2f ms; time-weighted tensor pipe active = %.1f %% of elapsed cycles" %
          (len(conv), ct, sum(l["gpu__time_duration.sum"] * l.get(key, 0.0) for l in conv) / ct))
    clk = [l["sm__cycles_elapsed.avg.per_second"] for l in conv if "sm__cycles_elapsed.avg.per_second" in l]
    if clk:
        print("mean SM clock during conv_tc launches: %.3g (ncu unit as reported)" % (sum(clk) / len(clk)))
    for lo, hi in ((90, 101), (80, 90), (60, 80), (40, 60), (0, 40)):
        t = sum(l["gpu__time_duration.sum"] for l in conv if lo <= l.get(key, 0.0) < hi)
        print("  launches at %3d-%3d %% tensor-active: %5.1f %% of conv time" % (lo, hi, 100 * t / ct))
