#!/usr/bin/env python3
"""Static SASS instruction mix of a kernel in libbpg.so (cuobjdump -sass): opcode counts of the whole function and of
its hottest loop (the largest backward-branch body).  Usage: sass_mix.py <kernel name regex> [libbpg.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(sys.argv[1])
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "bulletproof_gadgets_b200", "libbpg.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, body = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        body[fn] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m and fn:
        body[fn].append((int(m.group(1), 16), m.group(2)))
for fn, ins in body.items():
    if not pat.search(fn) or not ins:
        continue
    def opcode(t):
        toks = t.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        return op
    # hot loop = the smallest backward-branch body that gathers table rows (128-bit global loads); failing that, the largest
    loops = []
    for addr, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    gather = [l for l in loops if sum(1 for a, t in ins if l[0] <= a <= l[1] and "LDG.E.128" in t) >= 2]
    best = min(gather, key=lambda l: l[1] - l[0]) if gather else max(loops, key=lambda l: l[1] - l[0], default=(0, 0))
    for title, sel in (("whole function", ins), ("hot loop body 0x%x..0x%x (one mixed addition per trip)" % best, [(a, t) for a, t in ins if best[0] <= a <= best[1]])):
        c = collections.Counter(opcode(t) for _, t in sel)
        tot = sum(c.values())
        mul = sum(v for k, v in c.items() if k.startswith("IMAD"))
        print("%s -- %s: %d instructions (%.1f KB), %d on the multiplier pipe (IMAD*)" % (fn[:60], title, tot, tot * 16 / 1024, mul))
        for k, v in c.most_common(14):
            print("    %-22s %6d  %5.1f%%" % (k, v, 100.0 * v / tot))
