#!/usr/bin/env python
"""Decode scoreboard/stall control bits of a kernel's SASS (cuobjdump -sass): usage sassctl.py <lib.so> <function-substring> [grep]"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on = False; lines = []
for l in txt:
    if "Function :" in l:
        on = pat in l
        continue
    if on: lines.append(l)
i = 0; n = 0
while i < len(lines):
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
            stall = (w >> 105) & 0xf; wbar = (w >> 110) & 7; rbar = (w >> 113) & 7; wait = (w >> 116) & 0x3f
            t = m.group(2).strip()
            if flt is None or re.search(flt, t) or wait:
                print(f"{int(m.group(1),16):#07x} st={stall:2d} wb={'-' if wbar==7 else wbar} rb={'-' if rbar==7 else rbar} wait={wait:06b}  {t[:90]}")
            n += 1; i += 2; continue
    i += 1
print("instructions:", n)
