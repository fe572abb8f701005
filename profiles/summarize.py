#!/usr/bin/env python
"""Extract the numbers this repo quotes from an Nsight Compute report: python profiles/summarize.py <file.ncu-rep>"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")])
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:70s} {vals[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr = rows[1]
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot, ops, n = collections.Counter(), collections.Counter(), 0
        ia, ie = hdr.index("Source"), hdr.index("Instructions Executed")
        for r in rows[2:]:
            try:
                e = int(r[ie])
            except (ValueError, IndexError):
                continue
            n += e
            t = r[ia].split()
            ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += e
            for h in reasons:
                try:
                    tot[h] += int(r[hdr.index(h)])
                except ValueError:
                    pass
        s = sum(tot.values()) or 1
        print("  warp stall samples:", ", ".join(f"{h[6:]} {c / s * 100:.1f}%" for h, c in tot.most_common(8)))
        print("  executed SASS mix :", ", ".join(f"{o} {c / n * 100:.1f}%" for o, c in ops.most_common(12)))


if __name__ == "__main__":
    main(sys.argv[1])
