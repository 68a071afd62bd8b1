"""Turns an ncu --csv launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) into
(1) a compact per-launch table and (2) per-kernel totals / shares. usage: summarise_ncu.py in.csv out_prefix"""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
launches = collections.OrderedDict()
for r in rows:
    d = launches.setdefault(int(r[0]), {"kernel": r[4], "grid": r[8], "block": r[7]})
    v = float(r[14].replace(",", ""))
    unit = r[13]
    if r[12].startswith("gpu__time_duration"):
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("u") else v * 1e3)
    elif "read" in r[12]:
        d["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif "write" in r[12]:
        d["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("idf::", "")
    return re.sub(r"\(.*", "", name)[:60]


with open(dst + "_launches.csv", "w") as fh:
    fh.write("id,kernel,grid,block,duration_us,dram_read_MB,dram_write_MB\n")
    for i, d in launches.items():
        fh.write(f"{i},\"{short(d['kernel'])}\",\"{d['grid']}\",\"{d['block']}\",{d.get('us', 0):.2f},"
                 f"{d.get('rd', 0) / 1e6:.2f},{d.get('wr', 0) / 1e6:.2f}\n")
tot = collections.OrderedDict()
for d in launches.values():
    k = re.sub(r"<.*", "", short(d["kernel"]))
    t = tot.setdefault(k, {"launches": 0, "us": 0.0, "dram_MB": 0.0})
    t["launches"] += 1
    t["us"] += d.get("us", 0)
    t["dram_MB"] += (d.get("rd", 0) + d.get("wr", 0)) / 1e6
total_us = sum(t["us"] for t in tot.values())
for t in tot.values():
    t["share"] = round(t["us"] / total_us, 4)
    t["us"] = round(t["us"], 1)
    t["dram_MB_per_launch"] = round(t["dram_MB"] / t["launches"], 2)
    t["dram_MB"] = round(t["dram_MB"], 1)
out = {"source": src, "note": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control "
                              "none: per-launch times are cold-cache and serialised - the SHARES are what compares with "
                              "bench.py's CUDA-event breakdown", "total_us": round(total_us, 1),
       "kernels": dict(sorted(tot.items(), key=lambda kv: -kv[1]["us"]))}
json.dump(out, open(dst + "_shares.json", "w"), indent=1)
print(json.dumps(out["kernels"], indent=1))
