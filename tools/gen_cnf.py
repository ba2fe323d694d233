#!/usr/bin/env python3
"""Synthetic rsaFACT-Nbit factoring CNF generator (SURVEY.md §8c "probe generator").

The reference ships no inputs; what NDP *requires* of an input is: the four header
regexes (NDP-5_6_7.cpp:1591-1592,1598, 948-949), one clause per line, widths 1 or 3
only (NDP-5_6_7.cpp:632-639).  This generator emits an h x h array multiplier whose
every 2-input gate is 4 ternary truth-table clauses; product bits are unit clauses.

    python tools/gen_cnf.py N BITS out.dimacs      # BITS = total product bits (even)

Named corpus (fixed semiprimes / primes from SURVEY.md §8c):
    python tools/gen_cnf.py --named rsaFACT-32bit out.dimacs
"""
import sys

NAMED = {
    # name: (N, total product bits); factors have exactly bits/2 bits each
    "rsaFACT-8bit": (11 * 13, 8),                    # 143
    "rsaFACT-12bit": (47 * 53, 12),                  # 2491
    "rsaFACT-16bit": (35611, 16),                    # 149 * 239
    "rsaFACT-20bit": (751 * 997, 20),                # 748747
    "rsaFACT-24bit": (3323 * 3889, 24),              # 12923147
    "rsaFACT-32bit": (2500520291, 32),               # 45887 * 54493
    "rsaFACT-48bit": (209727620799133, 48),          # 15908531 * 13183343
    "rsaFACT-64bit": (11708736375489890129, 64),     # 3379623959 * 3464508631
    # primes with exactly `bits` bits: UNSAT as an h x h multiplier
    "prime-8bit": (251, 8),
    "prime-12bit": (4093, 12),
    "prime-16bit": (35051, 16),
    "prime-24bit": (14962111, 24),
    "prime-32bit": (4158910729, 32),
}


def multiplier_cnf(N, bits):
    """Return (num_vars, clauses, A_vars_lsb_first, B_vars_lsb_first)."""
    assert bits % 2 == 0 and bits >= 4
    h = bits // 2
    assert N >> bits == 0, "N does not fit in the product width"
    clauses = []
    nvars = 0

    def new_var():
        nonlocal nvars
        nvars += 1
        return nvars

    A = [new_var() for _ in range(h)]       # A[0] = lsb
    B = [new_var() for _ in range(h)]

    def gate(fn, x, y):
        z = new_var()
        for vx in (0, 1):
            for vy in (0, 1):
                clauses.append((-x if vx else x, -y if vy else y, z if fn(vx, vy) else -z))
        return z

    def AND(x, y):
        return gate(lambda p, q: p & q, x, y)

    def XOR(x, y):
        return gate(lambda p, q: p ^ q, x, y)

    def OR(x, y):
        return gate(lambda p, q: p | q, x, y)

    cols = [[] for _ in range(2 * h)]
    for i in range(h):
        for j in range(h):
            cols[i + j].append(AND(A[i], B[j]))
    outs = []
    for k in range(2 * h):
        col = cols[k]
        while len(col) > 1:
            if len(col) >= 3:
                x, y, c = col.pop(0), col.pop(0), col.pop(0)
                t = XOR(x, y)
                s = XOR(t, c)
                g1 = AND(x, y)
                g2 = AND(t, c)
                cout = OR(g1, g2)
            else:
                x, y = col.pop(0), col.pop(0)
                s = XOR(x, y)
                cout = AND(x, y)
            col.append(s)
            if k + 1 < 2 * h:
                cols[k + 1].append(cout)
            else:
                clauses.append((-cout,))
        outs.append(col[0] if col else None)
    for k, o in enumerate(outs):
        bit = (N >> k) & 1
        if o is None:
            assert bit == 0
            continue
        clauses.append((o if bit else -o,))
    return nvars, clauses, A, B


def dimacs_text(N, bits):
    nvars, clauses, A, B = multiplier_cnf(N, bits)
    lines = [
        "c Circuit for product = %d [%s]" % (N, bin(N)[2:]),
        "c Variables for first input [msb,...,lsb]: [%s]" % ", ".join(map(str, A[::-1])),
        "c Variables for second input [msb,...,lsb]: [%s]" % ", ".join(map(str, B[::-1])),
        "p cnf %d %d" % (nvars, len(clauses)),
    ]
    lines += [" ".join(map(str, c)) + " 0" for c in clauses]
    return "\n".join(lines) + "\n"


def write_named(name, path):
    N, bits = NAMED[name]
    with open(path, "w") as f:
        f.write(dimacs_text(N, bits))
    return N, bits


def main(argv):
    if len(argv) == 4 and argv[1] == "--named":
        print(write_named(argv[2], argv[3]))
        return 0
    if len(argv) != 4:
        sys.stderr.write(__doc__)
        return 2
    N, bits = int(argv[1]), int(argv[2])
    with open(argv[3], "w") as f:
        f.write(dimacs_text(N, bits))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
