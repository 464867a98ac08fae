#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` logs under build/: registers, spills, smem per kernel."""
import glob, re, subprocess, sys

def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n

rows = []
for f in sorted(glob.glob("build/*.ptxas.log")):
    txt = open(f).read().splitlines()
    cur = None
    for ln in txt:
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            cur = {"name": m.group(1), "spill": "0/0"}
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and cur is not None:
            cur["stack"] = m.group(1); cur["spill"] = f"{m.group(2)}/{m.group(3)}"
        m = re.search(r"Used (\d+) registers", ln)
        if m and cur is not None:
            cur["regs"] = m.group(1)
            sm = re.search(r"(\d+) bytes smem", ln)
            cur["smem"] = sm.group(1) if sm else "0"
            rows.append(cur); cur = None
for r in rows:
    d = demangle(r["name"])
    d = re.sub(r"fi::\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"{d:55s} regs={r['regs']:>3s} stack={r.get('stack','0'):>4s} spill={r['spill']:>9s} smem={r['smem']}")
