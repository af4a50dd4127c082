#!/usr/bin/env python3
"""Summarise an ncu report's source page per kernel: top SASS instructions by stall samples, grouped stall reasons.
usage: ncu_hot.py <report.ncu-rep> [kernel-substring] [topN]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name",')[1:]
for b in blocks:
    lines = b.splitlines()
    name = lines[0].strip('",')
    if sub not in name: continue
    rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rd[0]; ix = {h: i for i, h in enumerate(hdr)}
    rows = [r for r in rd[1:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]]) for r in rows)
    print("=====", name, "samples", tot, "instructions", len(rows))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in rows:
        for s in stalls: agg[s] += int(r[ix[s]] or 0)
    print("  stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in agg.most_common(8)))
    opc = collections.Counter(); opi = collections.Counter()
    for r in rows:
        op = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
        op = op.split(".")[0]
        opc[op] += int(r[ix["# Samples"]]); opi[op] += int(r[ix["Instructions Executed"]])
    toti = sum(opi.values())
    print("  by opcode (samples%, executed%):", ", ".join("%s %.1f/%.1f" % (k, 100.0 * v / max(tot, 1), 100.0 * opi[k] / max(toti, 1)) for k, v in opc.most_common(14)))
    rows_s = sorted(range(len(rows)), key=lambda i: -int(rows[i][ix["# Samples"]]))[:top]
    for i in sorted(rows_s):
        r = rows[i]
        st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print("  %5d %6s %10s  %-70s %s" % (i, r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70], " ".join("%s:%d" % (n, c) for c, n in st if c)))
