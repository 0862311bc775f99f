#!/usr/bin/env python3
"""Print registers / spill stack of every kernel in the built library (cuobjdump -res-usage)."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "ndrustfft_b200/lib/libndfft_b200.so"
out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
rows = []
name = None
for line in out.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1); continue
    m = re.search(r"REG:(\d+) STACK:(\d+)", line)
    if m and name:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("ndfb::", "").replace("void ", "")
        rows.append((dem, int(m.group(1)), int(m.group(2)))); name = None
for d, r, s in sorted(rows):
    flag = "  <-- SPILL" if s else ""
    print(f"REG={r:3d} STACK={s:4d}  {d[:150]}{flag}")
