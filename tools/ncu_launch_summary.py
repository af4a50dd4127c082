#!/usr/bin/env python3
"""Per kernel family summary of an `ncu --csv --metrics ...` launch list (tools/gpu_r2i.sh): launches, time, RED sectors,
warp instructions, DRAM bytes.  usage: ncu_launch_summary.py <launches.csv>"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None; L = collections.OrderedDict()
for r in rows:
    if len(r) > 10 and r[0] == "ID":
        hdr = r; continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = d.get("Metric Unit", "")
    if d["Metric Name"] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1e-6)
    if d["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    L.setdefault(d["ID"], {"name": d["Kernel Name"]})[d["Metric Name"]] = v
fam = collections.defaultdict(lambda: collections.defaultdict(float))
for v in L.values():
    f = re.match(r"(void )?(\w+)", v["name"]).group(2)
    for m, x in v.items():
        if m != "name":
            fam[f][m] += x
    fam[f]["n"] += 1
tot = sum(v["gpu__time_duration.sum"] for v in fam.values())
red = sum(v["l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum"] for v in fam.values())
print("%d launches, %.2f ms of kernel time under ncu (serialised, cold caches), %.3e RED sectors, DRAM %.1f MB read / %.1f MB written" % (
    len(L), tot, red, sum(v["dram__bytes_read.sum"] for v in fam.values()) / 1e6, sum(v["dram__bytes_write.sum"] for v in fam.values()) / 1e6))
for f, v in sorted(fam.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    print("  %-22s n=%4d  %9.3f ms (%5.1f%%)  RED sectors %.3e  warp instr %.3e  DRAM rd %8.1f MB wr %8.1f MB" % (
        f, v["n"], v["gpu__time_duration.sum"], 100 * v["gpu__time_duration.sum"] / tot, v["l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum"],
        v["smsp__inst_executed.sum"], v["dram__bytes_read.sum"] / 1e6, v["dram__bytes_write.sum"] / 1e6))
