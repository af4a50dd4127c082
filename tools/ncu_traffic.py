#!/usr/bin/env python3
"""Sum dram bytes / durations over all kernels of an ncu --csv log (tools/gpu_dram.sh) -> JSON entry for profiles/ncu_traffic.json.
usage: ncu_traffic.py <log.csv> <workload> <source note>"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; acc = {}; ids = set()
for r in rows:
    if len(r) > 10 and r[0] == "ID":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); ids.add(d["ID"])
        try:
            acc[d["Metric Name"]] = acc.get(d["Metric Name"], 0.0) + float(d["Metric Value"].replace(",", ""))
        except ValueError:
            pass
print(json.dumps({sys.argv[2]: {"dram_bytes_read": acc.get("dram__bytes_read.sum"), "dram_bytes_write": acc.get("dram__bytes_write.sum"),
                                "kernels": len(ids), "sum_kernel_ns_under_ncu": acc.get("gpu__time_duration.sum"), "source": sys.argv[3]}}, indent=1))
