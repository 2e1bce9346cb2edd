#!/usr/bin/env python
"""Turn an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) into the
per-kernel share table kept under profiles/ and the conv traffic figure bench.py reports as roofline.traffic.

    python tools/launch_summary.py profiles/r01_launches_v8.csv "title line" profiles/r01_launches_v8_summary.md \
        profiles/r01_conv_traffic.json
"""
import csv
import json
import re
import sys
from collections import OrderedDict


def main(src, title, md_out, traffic_out=None):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per_launch = OrderedDict()
    for r in rows[1:]:
        key = r[ix["ID"]]
        d = per_launch.setdefault(key, {"name": r[ix["Kernel Name"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
        d[r[ix["Metric Name"]]] = v * scale
    agg = OrderedDict()
    for d in per_launch.values():
        m = re.search(r"b2r::(\w+(?:<[^>]*>)?)", d["name"])
        if not m:
            continue
        a = agg.setdefault(m.group(1), {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(a["us"] for a in agg.values())
    lines = [f"# {title}", "",
             "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`; per-launch "
             "times are cold-cache and serialised: compare SHARES. `at::` kernels (checkpoint packing, buffer fills) are outside "
             "the timed region and omitted.", "",
             "| kernel | launches | total us | share of b2r time | DRAM read MB / launch | DRAM write MB / launch |",
             "|---|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        lines.append(f"| `b2r::{k}` | {a['n']} | {a['us']:.0f} | {100 * a['us'] / total:.1f}% | {a['rd'] / a['n'] / 1e6:.1f} | "
                     f"{a['wr'] / a['n'] / 1e6:.1f} |")
    open(md_out, "w").write("\n".join(lines) + "\n")
    if traffic_out:
        conv = [a for k, a in agg.items() if k.startswith("conv_")]
        n = sum(a["n"] for a in conv)
        rd, wr = sum(a["rd"] for a in conv), sum(a["wr"] for a in conv)
        json.dump({"source": src, "kernels": "all b2r::conv_* launches of one bench.py run under ncu", "launches": n,
                   "dram_bytes_per_launch": (rd + wr) / n, "dram_read_bytes_per_launch": rd / n,
                   "dram_write_bytes_per_launch": wr / n}, open(traffic_out, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:5])
