#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by CUDA source line.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> > sass.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g -c <kernel>.cubin > k.sass
  python tools/ncu_by_line.py sass.csv k.sass <mangled-kernel-substring> [top]

Prints executed warp-instructions and stall samples per file:line (innermost inlined location).
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, disasm, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    isrc = hdr.index("Source")
    per_addr = []
    end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] in ("Address", "Kernel Name")), len(rows))
    for r in rows[hi + 1:end]:
        if len(r) > ie and r[ie] not in ("", None):
            per_addr.append((r[isrc], float(r[ie] or 0), float(r[isamp] or 0)))
    # address -> line from nvdisasm (instruction order is the same as in the ncu page)
    lines, cur, on = [], ("?", 0), False
    for l in open(disasm):
        if l.startswith("//---") and ".text." in l:
            on = kname in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    n = min(len(lines), len(per_addr))
    if len(lines) != len(per_addr):
        print("warning: %d disassembled instructions vs %d profiled" % (len(lines), len(per_addr)), file=sys.stderr)
    agg = defaultdict(lambda: [0.0, 0.0, 0])
    for k in range(n):
        a = agg[lines[k]]
        a[0] += per_addr[k][1]
        a[1] += per_addr[k][2]
        a[2] += 1
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    print("total warp-instructions %.0f, stall samples %.0f, static instructions %d" % (tot, tots, n))
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-24s %5d  inst %12.0f (%5.1f%%)  samples %8.0f (%5.1f%%)  static %4d" % (f, ln, a[0], 100 * a[0] / tot, a[1], 100 * a[1] / tots, a[2]))


if __name__ == "__main__":
    main()
