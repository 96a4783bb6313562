"""Turns an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum --csv, one k_chain launch of `bench.py --steps 1
--warmup 1 --extra 0`) into web-audio-api-rs_b200/measured/dram_traffic.json, which bench.py reports as roofline.traffic.
usage: python tools/record_traffic.py ncu.csv graphs frames_per_graph out.json"""
import csv
import json
import os
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
h = rows[0]
ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tot, kernel = 0.0, None
first_id = rows[1][0]
for r in rows[1:]:
    if r[0] != first_id:  # first profiled launch only
        break
    kernel = r[ki]
    tot += float(r[vi].replace(",", "")) * scale[r[ui]]
out = {"kernel": "k_chain" if "k_chain" in kernel else kernel, "graphs": int(sys.argv[2]), "frames_per_graph": int(sys.argv[3]),
       "dram_bytes_per_launch": tot, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_chain -c 1 python bench.py --steps 1 --warmup 1 --extra 0"}
json.dump(out, open(sys.argv[4], "w"), indent=1)
print(out)
