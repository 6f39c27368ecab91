#!/usr/bin/env python3
"""Stall-reason totals and their top source lines for one kernel of an ncu report.
  ncu -i rep --page source --csv --kernel-name regex:<k> > sass.csv ; nvdisasm -g -c k.cubin > k.sass
  python tools/ncu_stalls.py sass.csv k.sass <mangled-substring>"""
import csv, re, sys
from collections import defaultdict
sass_csv, disasm, kname = sys.argv[1:4]
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; cols = {h: i for i, h in enumerate(hdr)}
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] in ("Address", "Kernel Name")), len(rows))
per = [r for r in rows[hi + 1:end] if len(r) > cols["Instructions Executed"] and r[cols["Instructions Executed"]] != ""]
lines, cur, on = [], ("?", 0), False
for l in open(disasm):
    if l.startswith("//---") and ".text." in l:
        on = kname in l; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(float(r[cols[h]] or 0) for r in per) for h in reasons}
s = sum(tot.values()) or 1
print("instructions", sum(float(r[cols["Instructions Executed"]] or 0) for r in per), "static", len(per), "samples", s)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]:
    print("%-26s %8.0f %5.1f%%" % (h, v, 100 * v / s))
    agg = defaultdict(float)
    for k in range(min(len(lines), len(per))): agg[lines[k]] += float(per[k][cols[h]] or 0)
    for (f, ln), x in sorted(agg.items(), key=lambda kv: -kv[1])[:4]:
        if x > 0: print("      %-22s %4d %7.0f %5.1f%%" % (f, ln, x, 100 * x / max(v, 1)))
