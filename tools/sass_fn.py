#!/usr/bin/env python3
"""Extract one kernel's SASS from a built library (no GPU needed):
    python tools/sass_fn.py ndp-5_b200/lib/libndp_b200.so dfs_index_kernel > /tmp/k.sass
Lines: <hex address> <instruction>.  Used to count the instructions on a kernel's hot path
before spending GPU time (B200_PROFILING.md: check cuobjdump -sass first)."""
import re
import subprocess
import sys


def main(lib, name):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
    on = False
    for line in out.splitlines():
        if "Function :" in line:
            on = name in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\*", line)
        if m:
            print(m.group(1), m.group(2))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
