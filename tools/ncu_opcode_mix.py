#!/usr/bin/env python3
"""Executed warp-instruction mix by opcode from the source page of an .ncu-rep (captured with --import-source on):
instructions executed per `unit` (e.g. per attempt = the execution count of the hottest straight-line region).
    python tools/ncu_opcode_mix.py file.ncu-rep"""
import collections
import csv
import io
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    k0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[k0], [r for r in rows[k0 + 1:] if len(r) > 10]
    ix = {h: i for i, h in enumerate(hdr)}
    print(rows[0][1] if len(rows[0]) > 1 else "")
    ex = [int(r[ix["Instructions Executed"]]) for r in data]
    unit = collections.Counter(e for e in ex if e).most_common(1)[0][0]        # (non-zero) execution count shared by the most instructions
    ops = collections.Counter()
    for r, e in zip(data, ex):
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        if m and e:
            ops[m.group(1)] += e
    tot = sum(ops.values())
    f64 = sum(v for k, v in ops.items() if k in FP64)
    print("unit = %d executions (the hottest region); %.1f warp instructions per unit, %.1f of them on the FP64 pipe (%s), %.1f other"
          % (unit, tot / unit, f64 / unit, "/".join(FP64), (tot - f64) / unit))
    print("issue-slot model (FP64 instruction = 2 slots): %.0f slots per unit; FP64 share %.1f %%" % ((tot + f64) / unit, 200.0 * f64 / (tot + f64)))
    for k, v in ops.most_common(28):
        print("  %-10s %8.2f per unit" % (k, v / unit))


if __name__ == "__main__":
    main()
