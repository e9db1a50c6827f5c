#!/usr/bin/env python3
"""Static code size by source line: counts SASS instructions per (file:line) of a kernel from `nvdisasm --print-line-info`.
Inlined frames are ignored (the innermost line owns the instruction).  usage: sass_static_lines.py <kernel substring> [top N]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    kernel = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    lib = os.environ.get("LYRA_B200_LIB", os.path.join(ROOT, "lyra_b200", "liblyra_b200.so"))
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, capture_output=True)
        cubins = [f for f in os.listdir(d) if f.endswith(".cubin")]
        cubin = max(cubins, key=lambda f: os.path.getsize(os.path.join(d, f)))
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    infn = False
    cur = ("?", 0)
    total = 0
    for line in txt.splitlines():
        if line.startswith(".text."):
            infn = kernel in line
            continue
        if not infn:
            continue
        if line.startswith(".section") or line.startswith(".text"):
            infn = False
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            if "inlined at" not in line:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            cnt[cur] += 1
            total += 1
    print(kernel, "instructions", total)
    byfile = collections.Counter()
    for (f, l), c in cnt.items():
        byfile[f] += c
    print(dict(byfile))
    for (f, l), c in cnt.most_common(top):
        print("%6d %5.1f%%  %s:%d" % (c, 100.0 * c / max(1, total), f, l))


if __name__ == "__main__":
    main()
