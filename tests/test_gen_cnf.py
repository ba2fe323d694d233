"""Input side: the generator's two header styles read identically through the reference's own parser
and header scrape, and the input check flags what the reference would silently mis-read."""
import numpy as np
import pytest

import cbind
import gen_cnf


@pytest.mark.skipif(not cbind.RefLib.available(), reason="the compiled reference did not travel here")
@pytest.mark.parametrize("name", ["rsaFACT-8bit", "rsaFACT-16bit", "prime-12bit"])
def test_header_styles_read_identically_by_the_reference(name):
    ref = cbind.RefLib()
    N, bits = gen_cnf.NAMED[name]
    a, b = gen_cnf.dimacs_text(N, bits, "minimal"), gen_cnf.dimacs_text(N, bits, "purdom-sabry")
    assert (ref.parse(a) == ref.parse(b)).all()
    ha, hb = ref.scrape_header(a), ref.scrape_header(b)
    for k in ("rc", "num_vars", "num_clauses", "num_bits", "product"):
        assert ha[k] == hb[k], k
    assert ha["v1"].tolist() == hb["v1"].tolist() and ha["v2"].tolist() == hb["v2"].tolist()
    assert int(ha["product"]) == N and ha["num_bits"] > 0


def test_oracle_parser_agrees_on_both_styles(oracle):
    N, bits = gen_cnf.NAMED["rsaFACT-12bit"]
    assert (oracle.parse(gen_cnf.dimacs_text(N, bits, "minimal")) == oracle.parse(gen_cnf.dimacs_text(N, bits, "purdom-sabry"))).all()


def test_input_check():
    N, bits = gen_cnf.NAMED["rsaFACT-8bit"]
    for style in ("minimal", "purdom-sabry"):
        assert gen_cnf.check_text(gen_cnf.dimacs_text(N, bits, style)) == []
    text = gen_cnf.dimacs_text(N, bits)
    lines = text.splitlines()
    bad = "\n".join(lines[:4] + ["1 -2 0"] + lines[4:]) + "\n"           # a width-2 clause: dropped by the reference
    p = gen_cnf.check_text(bad)
    assert any("width 2" in x and "DROPPED" in x for x in p)
    two = "\n".join(lines[:4] + ["1 2 3 0 4 5 6 0"] + lines[5:]) + "\n"  # two clauses on one line
    assert any("ONE clause per line" in x for x in gen_cnf.check_text(two))
    assert any("Circuit for product" in x for x in gen_cnf.check_text("p cnf 3 1\n1 2 3 0\n"))
    assert any("beyond the declared" in x for x in gen_cnf.check_text(text.replace("p cnf", "p cnf 2 0\nc p cnf", 1)))
