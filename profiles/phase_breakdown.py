#!/usr/bin/env python
"""Where a kernel's warp-stall samples fall, from an Nsight Compute report (needs -lineinfo and --import-source on).
  python profiles/phase_breakdown.py <file.ncu-rep> phases     # SASS order cut at block barriers / mbarrier waits (K1: pass A | B+C | output | tail)
  python profiles/phase_breakdown.py <file.ncu-rep> lines [N]  # the N source lines with the most samples (+ their share of executed instructions)"""
import collections
import csv
import subprocess
import sys


def page(path, view):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def phases(path):
    rows = page(path, "sass")
    print(rows[0][1])
    hdr, data = rows[1], rows[2:]
    ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    cur = {"n": 0, "samples": 0, "exec": 0, "st": collections.Counter()}
    segs, tot = [], 0
    for i, r in enumerate(data):
        try:
            s, e = int(r[isamp]), int(r[iex])
        except (ValueError, IndexError):
            continue
        cur["n"] += 1
        cur["samples"] += s
        cur["exec"] += e
        tot += s
        for h in stalls:
            try:
                cur["st"][h] += int(r[hdr.index(h)])
            except ValueError:
                pass
        t = r[ia].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        if op == "BAR" or ("SYNCS" in r[ia] and "TRYWAIT" in r[ia]):
            segs.append((i, r[ia].strip(), cur))
            cur = {"n": 0, "samples": 0, "exec": 0, "st": collections.Counter()}
    segs.append((len(data), "END", cur))
    for i, src, c in segs:
        top = " ".join(f"{k[6:]}={v}" for k, v in c["st"].most_common(5))
        print(f"up to SASS #{i:5d} {src[:44]:44s} static instr {c['n']:5d}  samples {c['samples']:6d} ({100 * c['samples'] / max(tot, 1):5.1f} %)  executed {c['exec']:10d}  {top}")
    print("total samples", tot)


def lines(path, top_n):
    rows = page(path, "cuda,sass")
    per, cur_file, hdr = collections.OrderedDict(), None, None
    for r in rows:
        if r and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) < 6 or r[0] == "":
            continue
        try:
            ln, s = int(r[0]), int(r[4])
        except ValueError:
            continue
        iex = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
        ex = int(r[iex]) if iex is not None and r[iex].isdigit() else 0
        a = per.setdefault((cur_file, ln), [0, 0, r[1].strip()[:100]])
        a[0] += s
        a[1] += ex
    tot, totex = sum(v[0] for v in per.values()) or 1, sum(v[1] for v in per.values()) or 1
    print(f"total samples {tot}, executed warp instructions {totex}")
    for (f, ln), v in sorted(per.items(), key=lambda x: -x[1][0])[:top_n]:
        print(f"{f}:{ln:4d}  samples {100 * v[0] / tot:5.1f} %  instructions {100 * v[1] / totex:5.1f} %   {v[2]}")


if __name__ == "__main__":
    if sys.argv[2] == "phases":
        phases(sys.argv[1])
    else:
        lines(sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
