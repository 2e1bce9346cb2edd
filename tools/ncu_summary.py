#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small markdown table for profiles/."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
]


def main(rep, title):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {title}\n")
    print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`), {len(data)} launches.\n")
    names = [r[hdr.index("Kernel Name")][:60] for r in data]
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|---|" + "---:|" * len(data))
    print("| kernel | | " + " | ".join(names) + " |")
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w or h.endswith("." + w):
                print(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
                break


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
