#!/usr/bin/env python3
"""Aggregate `ncu --page source --print-source sass,cuda --csv` by CUDA source line: stall samples and instructions per file:line.
usage: ncu_source_lines.py <rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kernel],
                                  stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    agg = {}
    cur_file, hdr, ix = "?", None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        try:
            line = int(r[0])
            samples = int(float(r[ix["# Samples"]] or 0))
            inst = int(float(r[ix["Instructions Executed"]] or 0))
        except ValueError:
            continue
        key = (cur_file, line)
        a = agg.setdefault(key, [0, 0, r[1].strip(), {}])
        a[0] += samples
        a[1] += inst
        for h, i in ix.items():
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    v = int(float(r[i] or 0))
                except ValueError:
                    v = 0
                if v:
                    a[3][h[6:]] = a[3].get(h[6:], 0) + v
    tot = sum(v[0] for v in agg.values())
    tinst = sum(v[1] for v in agg.values())
    print("kernel", kernel, "stall samples", tot, "warp instructions", tinst)
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        why = " ".join("%s=%d" % kv for kv in sorted(v[3].items(), key=lambda kv: -kv[1])[:3])
        print("%6d %5.1f%%  inst %5.1f%%  %s:%d  [%s]  %s" % (v[0], 100.0 * v[0] / max(1, tot), 100.0 * v[1] / max(1, tinst), f, l, why, v[2][:90]))


if __name__ == "__main__":
    main()
