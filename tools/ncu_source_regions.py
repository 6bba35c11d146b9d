#!/usr/bin/env python3
"""Summarise the source page of an .ncu-rep (captured with --import-source on): contiguous SASS ranges with the
same execution count / active-thread average, their share of all issued warp instructions and the dominant
stall reasons.  Usage: python tools/ncu_source_regions.py file.ncu-rep [min_share_percent]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    k0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[k0], [r for r in rows[k0 + 1:] if len(r) > 10]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    thr = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
    samples = sum(int(r[ix["# Samples"]]) for r in data)
    print("warp instructions %d, average active threads %.2f, stall samples %d" % (tot, thr / max(tot, 1), samples))
    runs = []
    for k, r in enumerate(data):
        ie, at = int(r[ix["Instructions Executed"]]), float(r[ix["Avg. Threads Executed"]])
        if runs and runs[-1][2] == ie and abs(runs[-1][3] - at) < 0.01:
            runs[-1][1] = k
        else:
            runs.append([k, k, ie, at])
    for a, b, ie, at in runs:
        n = b - a + 1
        share = 100.0 * n * ie / max(tot, 1)
        if share < min_share:
            continue
        st = {c: sum(int(data[k][ix[c]]) for k in range(a, b + 1)) for c in stall_cols}
        ns = sum(int(data[k][ix["# Samples"]]) for k in range(a, b + 1))
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print("SASS %4d..%4d (%3d instr) exec %10d  thr %5.1f  issue share %5.2f %%  samples %5.2f %%  %s" % (
            a, b, n, ie, at, share, 100.0 * ns / max(samples, 1),
            " ".join("%s=%.0f%%" % (c[6:], 100.0 * v / max(ns, 1)) for c, v in top)))


if __name__ == "__main__":
    main()
