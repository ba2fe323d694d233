"""Small helpers shared by the parity tests."""
import hashlib

import numpy as np


def frontier_digest(frontier):
    """Same digest as tools/make_golden.py: sha256 over (n, m, clauses, choices) in queue order."""
    h = hashlib.sha256()
    for cl, ch in frontier:
        cl = np.asarray(cl, dtype=np.int32).reshape(-1, 3)
        ch = np.asarray(ch, dtype=np.int32).reshape(-1)
        h.update(np.asarray([cl.shape[0], ch.shape[0]], dtype="<i4").tobytes())
        h.update(np.ascontiguousarray(cl, dtype="<i4").tobytes())
        h.update(np.ascontiguousarray(ch, dtype="<i4").tobytes())
    return h.hexdigest()


def element_digest(cl, ch):
    return frontier_digest([(cl, ch)])[:16]


def random_cnf(rng, n_vars, n_clauses, p_unit=0.1, p_binary=0.1, p_dup=0.05, p_taut=0.05, max_var=None):
    """Random n x 3 clause array in the reference's in-memory form: unit clauses as {0,0,x},
    pre-zeroed literals ({0,x,y} in any position), duplicate literals and tautologies."""
    out = np.zeros((n_clauses, 3), dtype=np.int32)
    hi = max_var or n_vars
    for k in range(n_clauses):
        vs = rng.choice(np.arange(1, hi + 1), size=3, replace=hi < 3)
        if hi != n_vars:  # sparse variable ids
            vs = rng.integers(1, hi + 1, size=3)
        lits = [int(v) * (1 if rng.random() < 0.5 else -1) for v in vs]
        r = rng.random()
        if r < p_unit:
            lits = [0, 0, lits[2]]
        elif r < p_unit + p_binary:
            z = int(rng.integers(0, 3))
            lits[z] = 0
        elif r < p_unit + p_binary + p_dup:
            lits[1] = lits[0]
        elif r < p_unit + p_binary + p_dup + p_taut:
            lits[2] = -lits[0]
        out[k] = lits
    return out


def settings_for(mode, value):
    """(max_iterations, override, max_queues) as main() passes them (NDP:1511-1512,1521,1651)."""
    return (value, False, -1) if mode == "d" else (0, True, value)


def decode_factors(model, v1, v2):
    """processVector/convert (NDP:1008-1027): bit = 1 iff the positive literal is in the model."""
    s = set(int(x) for x in model)
    d1 = int("".join("1" if int(v) in s else "0" for v in v1) or "0", 2)
    d2 = int("".join("1" if int(v) in s else "0" for v in v2) or "0", 2)
    return d1, d2


def satisfies(clauses, model):
    """Every original clause has a literal that the model makes true."""
    s = set(int(x) for x in model)
    cl = np.asarray(clauses).reshape(-1, 3)
    for a, b, c in cl:
        if not ((a != 0 and int(a) in s) or (b != 0 and int(b) in s) or (c != 0 and int(c) in s)):
            return False
    return True
