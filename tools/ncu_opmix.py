#!/usr/bin/env python3
"""Executed-instruction mix per kernel from an .ncu-rep source page (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = None; hdr = None; seen = set()
mix = collections.Counter(); tot = 0
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "Kernel Name":
        if kern and kern not in seen and (pat in kern):
            seen.add(kern)
            print("==", kern[:150]); print("   total warp-instr executed:", tot)
            for op, c in mix.most_common(22): print(f"   {c:12d} {100*c/max(tot,1):5.1f}%  {op}")
        kern = row[1]; hdr = None; mix = collections.Counter(); tot = 0
        continue
    if row[0] == "Address": hdr = row; continue
    if hdr is None: continue
    src = row[1].strip(); ex = int(row[hdr.index("Instructions Executed")] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0].rstrip(";")
    mix[op] += ex; tot += ex
if kern and kern not in seen and (pat in kern):
    print("==", kern[:150]); print("   total warp-instr executed:", tot)
    for op, c in mix.most_common(22): print(f"   {c:12d} {100*c/max(tot,1):5.1f}%  {op}")
