#!/usr/bin/env python3
"""Aggregate `ncu --page source --csv` output: opcode mix, shared-memory wavefronts per instruction, hottest SASS lines."""
import csv
import subprocess
import sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                                  stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]

    def num(r, k):
        try:
            return int(float(r[ix[k]]))
        except (ValueError, KeyError):
            return 0
    tot = sum(num(r, "# Samples") for r in body)
    print("kernel", kernel, "SASS lines", len(body), "stall samples", tot)
    agg = {}
    for r in body:
        src = r[ix["Source"]].strip()
        op = src.split()[0] if src else ""
        if op.startswith("@"):
            op = src.split()[1]
        a = agg.setdefault(op, [0, 0, 0])
        a[0] += num(r, "Instructions Executed")
        a[1] += num(r, "L1 Wavefronts Shared")
        a[2] += num(r, "# Samples")
    tinst = sum(v[0] for v in agg.values())
    print("%-28s %12s %6s %12s %8s %8s" % ("opcode", "warp-instr", "%", "smem wavefr", "wf/inst", "samples%"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
        print("%-28s %12d %6.1f %12d %8.2f %8.1f" % (k, v[0], 100.0 * v[0] / tinst, v[1], v[1] / max(1, v[0]), 100.0 * v[2] / max(1, tot)))
    print("--- hottest SASS lines by stall samples")
    for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
        print("%7d %5.1f%%  exec %9d  %s" % (num(r, "# Samples"), 100.0 * num(r, "# Samples") / max(1, tot), num(r, "Instructions Executed"), r[ix["Source"]].strip()[:100]))


if __name__ == "__main__":
    main()
