#!/usr/bin/env python3
"""Static SASS summary of the built library (no GPU needed): per kernel, the instruction count and the mix of the
mnemonics that matter for this path -- warp votes / shuffles / reductions (ballot-based conflict and unit detection),
global / shared / atomic traffic, control-flow reconvergence (BSSY / BSYNC), barriers -- and the proof that no tensor-core
or TMA instruction is present (none is wanted: the path has no dense contraction).

    python tools/sass_histogram.py [ndp-5_b200/lib/libndp_b200.so] > profiles/r2_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

GROUPS = [
    ("votes/shuffles/redux", ("VOTE", "VOTEU", "SHFL", "REDUX", "MATCH")),
    ("popc/flo/lop", ("POPC", "FLO", "LOP3", "BREV", "SHF")),
    ("global ld", ("LDG",)), ("global st", ("STG",)), ("shared ld", ("LDS",)), ("shared st", ("STS",)),
    ("const ld", ("LDC", "ULDC", "LDCU")), ("local ld/st (stack frame + spills)", ("LDL", "STL")),
    ("atomics", ("ATOM", "ATOMG", "ATOMS", "RED")), ("reconvergence", ("BSSY", "BSYNC")), ("branches", ("BRA", "BRX", "EXIT", "RET", "CALL")),
    ("barriers", ("BAR", "WARPSYNC", "MEMBAR", "ERRBAR", "CCTL")),
    ("tensor core / TMA", ("HMMA", "IMMA", "QMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UBLKCP", "TCGEN05", "UTCBAR", "LDTM", "STTM")),
]


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and name:
            kernels[name][m.group(1).split(".")[0]] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print("# static SASS summary of %s (cuobjdump -sass; counts are static instructions, not executed ones)" % lib)
    total_tc = 0
    for (mangled, ctr), pretty in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", pretty).replace("ndp::", "").replace("void ", "")
        n = sum(ctr.values())
        parts = []
        for label, names in GROUPS:
            c = sum(v for k, v in ctr.items() if k in names)
            if label == "tensor core / TMA":
                total_tc += c
            if c:
                parts.append("%s %d" % (label, c))
        print("%-58s %6d instr | %s" % (short, n, ", ".join(parts)))
    print("# tensor-core / TMA / tcgen05 instructions in the whole library: %d" % total_tc)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "ndp-5_b200/lib/libndp_b200.so")
