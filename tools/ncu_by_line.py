#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump (SASS level) by CUDA source line.

usage: ncu_by_line.py <src.csv from ncu -i rep --page source --csv --kernel-name regex:K> <nvdisasm -g -c dump> <function substring> [top [result]]
A report with several captured launches prints one table after the other: `result` (default 0) picks the table.
The n-th SASS instruction of the ncu dump is matched with the n-th instruction of the nvdisasm
listing of the same function (identical binaries), whose `//## File "...", line N` markers give the line.
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, func):
    lines = open(path).read().split("\n")
    out, cur, infn = [], None, False
    for ln in lines:
        if ln.startswith(".text."):
            infn = func in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            # inlined-at info: keep the innermost
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append((cur, ln.strip()))
    return out


def main():
    src, dis, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    idx = {n: i for i, n in enumerate(hdr)}
    tables, cur = [], None
    for r in rows[1:]:
        if r == hdr:                                   # every captured launch repeats the header line
            cur = []
            tables.append(cur)
        elif cur is not None and len(r) > 10:
            cur.append(r)
    data = tables[which]
    sl = sass_lines(dis, func)
    if len(sl) != len(data):
        print(f"warning: {len(data)} ncu instructions vs {len(sl)} nvdisasm instructions", file=sys.stderr)
    agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    ti = ts = 0.0
    for i, r in enumerate(data):
        key = sl[i][0] if i < len(sl) else None
        ins = float(r[idx["Instructions Executed"]] or 0)
        smp = float(r[idx["# Samples"]] or 0)
        agg[key][0] += smp
        agg[key][1] += ins
        for c in stall_cols:
            v = float(r[idx[c]] or 0)
            if v:
                agg[key][2][c] += v
        ti += ins
        ts += smp
    print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
    for key, (smp, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join(f"{k[6:]}={v / max(smp, 1) * 100:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{smp / ts * 100:5.1f}% smp {ins / ti * 100:5.1f}% ins  {key}  [{tops}]")


if __name__ == "__main__":
    main()
