#!/usr/bin/env python3
"""Synthetic rsaFACT-Nbit factoring CNF generator (SURVEY.md §8c "probe generator").

The reference ships no inputs; what NDP *requires* of an input is: the four header
regexes (NDP-5_6_7.cpp:1591-1592,1598, 948-949), one clause per line, widths 1 or 3
only (NDP-5_6_7.cpp:632-639).  This generator emits an h x h array multiplier whose
every 2-input gate is 4 ternary truth-table clauses; product bits are unit clauses.

    python tools/gen_cnf.py N BITS out.dimacs      # BITS = total product bits (even)

Named corpus (fixed semiprimes / primes from SURVEY.md §8c):
    python tools/gen_cnf.py --named rsaFACT-32bit out.dimacs [minimal|purdom-sabry]
Input check (what the reference requires of ANY generator's file, README.md:18-26):
    python tools/gen_cnf.py --check some.dimacs
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


def dimacs_text(N, bits, header="minimal"):
    """header="minimal": the four lines the reference's regexes need (NDP-5_6_7.cpp:1591-1592,1598, 948-949).
    header="purdom-sabry": the fuller comment block of the generators the reference's README names
    (README.md:18-26: Purdom & Sabry's CNF generator, GridSAT/CNF_FACT-MULT) -- a provenance line, the product line,
    an output-variables line, the two input lines, blank comment lines.  Neither generator is available offline, so the
    block is reconstructed from the reference's own regexes and README; what is CHECKED (tests/test_gen_cnf.py) is that
    the reference's parser and header scrape read both variants identically."""
    nvars, clauses, A, B = multiplier_cnf(N, bits)
    product = "c Circuit for product = %d [%s]" % (N, bin(N)[2:])
    first = "c Variables for first input [msb,...,lsb]: [%s]" % ", ".join(map(str, A[::-1]))
    second = "c Variables for second input [msb,...,lsb]: [%s]" % ", ".join(map(str, B[::-1]))
    problem = "p cnf %d %d" % (nvars, len(clauses))
    if header == "minimal":
        lines = [product, first, second, problem]
    elif header == "purdom-sabry":
        outs = [abs(c[0]) for c in clauses[-bits:]] if bits <= len(clauses) else []
        lines = ["c Problem generated from an array multiplication circuit (tools/gen_cnf.py, Purdom-Sabry style header)",
                 "c", product, "c Variables for output [msb,...,lsb]: [%s]" % ", ".join(map(str, outs[::-1])), first, second,
                 "c", "c", problem]
    else:
        raise ValueError("unknown header style %r" % header)
    lines += [" ".join(map(str, c)) + " 0" for c in clauses]
    return "\n".join(lines) + "\n"


def check_text(text):
    """What the reference REQUIRES of an input file, checked the way it reads one (NDP-5_6_7.cpp:619-642 parser,
    1590-1615 / 946-990 header scrape).  Returns a list of problem strings (empty = the file is usable):
    missing header regexes, clause lines whose width is not 1 or 3 (the reference silently drops them, NDP:639),
    several clauses on one line, literals beyond the declared variable count."""
    import re
    problems = []
    m = re.search(r"p cnf ([0-9]+) ([0-9]+)", text)
    if not m:
        problems.append("no `p cnf V C` line (NDP:1592)")
    if not re.search(r"Circuit for product = ([0-9]+) \[", text):
        problems.append("no `Circuit for product = N [` line (NDP:1591)")
    if not re.search(r"Variables for first input \[msb,...,lsb\]: \[(.*?)\]", text):
        problems.append("no `Variables for first input [msb,...,lsb]: [...]` line (NDP:948)")
    if not re.search(r"Variables for second input \[msb,...,lsb\]: \[.*?,\s*(\d+)\]", text):
        problems.append("no `Variables for second input [msb,...,lsb]: [..., v]` line (NDP:949,1598)")
    nvars = int(m.group(1)) if m else None
    n_declared = int(m.group(2)) if m else None
    n_ok = 0
    for ln, line in enumerate(text.splitlines(), 1):
        t = line.strip()
        if not t or t[0] in "cp%":
            continue
        toks = t.split()
        try:
            lits = [int(x) for x in toks]
        except ValueError:
            problems.append("line %d: not a clause: %r" % (ln, t[:40]))
            continue
        if lits.count(0) > 1 or (lits and lits[-1] != 0):
            problems.append("line %d: the reference reads ONE clause per line, terminated by 0" % ln)
            continue
        body = lits[:-1]
        if len(body) not in (1, 3):
            problems.append("line %d: clause of width %d is silently DROPPED by the reference (NDP:639); only 1 and 3 are read"
                            % (ln, len(body)))
            continue
        if nvars is not None and any(abs(l) > nvars for l in body):
            problems.append("line %d: literal beyond the declared %d variables" % (ln, nvars))
        n_ok += 1
    if n_declared is not None and n_ok != n_declared:
        problems.append("`p cnf` declares %d clauses, %d usable clause lines found" % (n_declared, n_ok))
    return problems


def write_named(name, path, header="minimal"):
    N, bits = NAMED[name]
    with open(path, "w") as f:
        f.write(dimacs_text(N, bits, header))
    return N, bits


def main(argv):
    if len(argv) == 4 and argv[1] == "--named":
        print(write_named(argv[2], argv[3]))
        return 0
    if len(argv) == 5 and argv[1] == "--named" and argv[4] in ("minimal", "purdom-sabry"):
        print(write_named(argv[2], argv[3], argv[4]))
        return 0
    if len(argv) == 3 and argv[1] == "--check":
        with open(argv[2], errors="replace") as f:
            problems = check_text(f.read())
        for p in problems:
            print(p)
        print("%s: %s" % (argv[2], "usable by NDP as it is" if not problems else "%d problem(s)" % len(problems)))
        return 1 if problems else 0
    if len(argv) != 4:
        sys.stderr.write(__doc__)
        return 2
    N, bits = int(argv[1]), int(argv[2])
    with open(argv[3], "w") as f:
        f.write(dimacs_text(N, bits))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
