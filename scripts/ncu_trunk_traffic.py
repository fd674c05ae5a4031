"""Aggregates an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`scripts/prof_emb.py emb N` (N = one embedding sub-batch, 296 by default; two passes, the second is used) into
profiles/rNN_trunk_traffic.json, which bench.py reads for `roofline.traffic`.
Usage: python scripts/ncu_trunk_traffic.py launches.csv out.json [segments]"""
import collections
import csv
import json
import re
import sys

TRUNK = ("conv1_kernel", "conv_block32_kernel", "conv_block64_kernel", "conv_tc_kernel", "conv_tc4_kernel",
         "conv_tc3_kernel", "conv_tc5_kernel", "conv_s2_kernel")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main(src, dst, segments=296):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        name = re.sub(r"<.*", "", r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", ""))
        d = per.setdefault(int(r[ix["ID"]]), {"name": name})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT[r[ix["Metric Unit"]]]
    launches = [d for _, d in sorted(per.items()) if d["name"] in TRUNK]
    half = launches[len(launches) // 2:]                     # second (warm) pass
    agg = collections.OrderedDict()
    for d in half:
        a = agg.setdefault(d["name"], {"launches": 0, "ms": 0.0, "dram_read_gb": 0.0, "dram_write_gb": 0.0})
        a["launches"] += 1
        a["ms"] += d["gpu__time_duration.sum"] * 1e3
        a["dram_read_gb"] += d["dram__bytes_read.sum"] / 1e9
        a["dram_write_gb"] += d["dram__bytes_write.sum"] / 1e9
    total = sum(a["dram_read_gb"] + a["dram_write_gb"] for a in agg.values()) * 1e9
    out = {"source": src, "segments": segments, "launches_per_pass": len(half),
           "ms_per_pass_under_ncu": sum(a["ms"] for a in agg.values()), "dram_bytes_per_pass": total,
           "per_kernel": agg}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 296)
