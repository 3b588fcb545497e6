#!/usr/bin/env python
"""Print the SASS of one kernel between two addresses.  usage: sass_range.py <lib.so> <mangled-name substring> <lo hex> <hi hex>"""
import re, subprocess, sys
lib, key, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
on = False
for l in subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines():
    if "Function :" in l:
        on = key in l
        continue
    if not on:
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and lo <= int(m.group(1), 16) <= hi:
        print(f"{int(m.group(1), 16):05x}  {m.group(2).strip()}")
