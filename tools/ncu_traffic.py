#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
`bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary` into profiles/<round>_traffic.json (what bench.py reads for
roofline.traffic).  usage: tools/ncu_traffic.py launches.csv out.json blocks [full_raw.csv]"""
import csv, json, sys

launches, out, blocks = sys.argv[1], sys.argv[2], int(sys.argv[3])
full = sys.argv[4] if len(sys.argv) > 4 else None
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
h = rows[0]
ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
per = {}
for r in rows[1:]:
    d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
kern = {}
for key in ("k_resample_rx_v3", "k_detect_design", "k_equalize_ring"):
    ls = [d for d in per.values() if key in d["name"]]
    big = max(d["gpu__time_duration.sum"] for d in ls)
    ls = [d for d in ls if d["gpu__time_duration.sum"] > 0.5 * big]          # the full-size launches of the step
    n = len(ls)
    kern[key] = {"launches": n,
                 "dram_read_bytes": sum(d["dram__bytes_read.sum"] for d in ls) / n,
                 "dram_write_bytes": sum(d["dram__bytes_write.sum"] for d in ls) / n,
                 "ncu_duration_ns": sum(d["gpu__time_duration.sum"] for d in ls) / n}
    kern[key]["dram_bytes"] = kern[key]["dram_read_bytes"] + kern[key]["dram_write_bytes"]
if full:
    rows = list(csv.reader(open(full)))
    h = rows[0]
    ki = h.index("Kernel Name")
    for key in kern:
        for r in rows[2:]:
            if key in r[ki]:
                d = dict(zip(h, r))
                kern[key]["issue_active_pct"] = float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"])
                kern[key]["fma_pipe_active_pct"] = float(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"])
                break
json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum per launch (mean over the full-size "
                   "launches of the three kernels) from ncu (--clock-control none) on `python bench.py --steps 2 --warmup 3 "
                   "--no-e2e --no-cpu-baseline --no-secondary`; raw list: %s%s" % (launches.split("/")[-1], "; issue/fma pipe pcts from " + full.split("/")[-1] if full else ""),
           "blocks": blocks, "kernels": kern}, open(out, "w"), indent=1)
print(json.dumps(kern, indent=1))
