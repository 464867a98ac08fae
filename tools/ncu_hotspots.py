#!/usr/bin/env python
"""Source-level hot spots of every kernel in an .ncu-rep (captured with --import-source on, built -lineinfo):
per kernel the stall-reason mix and the source lines with the most warp-stall samples / executed instructions.
usage: ncu_hotspots.py <report.ncu-rep> [requests_per_launch]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
names = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
kernels = []
for r in list(csv.reader(io.StringIO(names)))[2:]:
    d = r[4] if len(r) > 4 else ""
    if d and d not in kernels:
        kernels.append(d)
for idx in range(len(kernels)):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, agg, func = None, None, [], "?"
    stall = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) > 8:
            if r[2] == "-":
                try:
                    agg.append((cur, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0)))
                except ValueError:
                    pass
    # stall-reason mix from the SASS-only page (its columns line up with the header)
    out2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                          capture_output=True, text=True).stdout
    r2 = list(csv.reader(io.StringIO(out2)))
    stall = {}
    if len(r2) > 2:
        h2 = r2[1]
        cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
        for r in r2[2:]:
            for i in cols:
                try:
                    stall[h2[i]] = stall.get(h2[i], 0) + int(r[i])
                except (ValueError, IndexError):
                    pass
    tot_i = sum(a[4] for a in agg) or 1
    tot_s = sum(a[3] for a in agg) or 1
    print("=" * 110)
    print(func[:108])
    print(f"warp instructions executed: {tot_i}" + (f"  ({tot_i / per:.0f} per request)" if per else "") + f"; stall samples: {tot_s}")
    T = sum(stall.values()) or 1
    print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / T:.1f}%" for k, v in sorted(stall.items(), key=lambda x: -x[1]) if 100 * v >= T))
    print("-- lines by stall samples")
    for f, ln, src, s_, i in sorted(agg, key=lambda a: -a[3])[:12]:
        print(f"  {100 * s_ / tot_s:5.1f}% samples  {100 * i / tot_i:5.1f}% instr  {f}:{ln}  {src[:70]}")
    print("-- lines by executed instructions")
    for f, ln, src, s_, i in sorted(agg, key=lambda a: -a[4])[:10]:
        print(f"  {100 * i / tot_i:5.1f}% instr  {100 * s_ / tot_s:5.1f}% samples  {f}:{ln}  {src[:70]}")
