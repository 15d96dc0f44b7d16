#!/usr/bin/env python3
"""Summarises `ncu -i <rep> --page source --csv --kernel-name regex:<k>` (stdin or file): stall reasons over all sampled
warps and the instruction mix / sample share per SASS opcode.  Usage: ncu_source_summary.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
col = {name: i for i, name in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
samples, byop, insts = 0, {}, {}
for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    n = int(r[col["# Samples"]] or 0)
    samples += n
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
    toks = r[col["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    base = op.split(".")[0] + (".WIDE" if "WIDE" in op else "") + (".HI" if ".HI" in op else "")
    byop[base] = byop.get(base, 0) + n
    insts[base] = insts.get(base, 0) + int(r[col["Instructions Executed"]] or 0)
print("kernel:", rows[0][1][:100] if rows[0] else "")
print("sampled warps:", samples)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-26s %8d %5.1f%%" % (s, v, 100.0 * v / max(samples, 1)))
ti = sum(insts.values())
print("warp instructions executed:", ti)
for o, v in sorted(insts.items(), key=lambda kv: -kv[1])[:16]:
    print("  %-12s insts %10d %5.1f%%   samples %8d %5.1f%%" % (o, v, 100.0 * v / ti, byop[o], 100.0 * byop[o] / max(samples, 1)))
