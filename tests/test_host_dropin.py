"""The C++17 host drop-in (ndp-5_b200/host) against the reference's own functions: DIMACS parser,
header scrape, CLI, naming, formatting, GMP factor check, report text, BFS JSON in both
directions (our writer -> reference reader, reference writer -> our reader).  CPU only.
Known answers (SURVEY.md §8c "formats") pin the behaviour where the reference build is absent."""
import json
import os
import re

import numpy as np
import pytest

import cbind
import util_cnf

HAVE_REF = cbind.RefLib.available()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    return cbind.HostLib()


@pytest.fixture(scope="module")
def ref():
    if not HAVE_REF:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return cbind.RefLib()


NASTY = "c comment\np cnf 5 6\n1 -2 3 0\n\n-4 0\n1 2 0\n1 2 3 4 0\n 5 0\nx 1 2 3 0\n+6 -7 8 0\r\n\r\n" \
        "99999999999 1 2 0\n1 2 99999999999 0\n12abc 3 4 0\n- 1 2 3 0\n2 -3 4\n7 8 9 0 10 11 0\n%\n0\n2 -3 4"


def test_parser_known_answers(host, oracle, cnf_text):
    assert host.parse(NASTY).tolist() == oracle.parse(NASTY).tolist()
    assert host.parse("").shape == (0, 3)
    for name in ("rsaFACT-8bit", "rsaFACT-16bit", "prime-16bit"):
        a, b = host.parse(cnf_text(name)), oracle.parse(cnf_text(name))
        assert a.shape == b.shape and (a == b).all()
    assert host.parse(cnf_text("rsaFACT-16bit")).shape == (1296, 3)


def test_parser_and_header_match_reference(host, ref, cnf_text):
    for text in [NASTY, cnf_text("rsaFACT-16bit"), cnf_text("rsaFACT-32bit")]:
        a, b = host.parse(text), ref.parse(text)
        assert a.shape == b.shape and (a == b).all()
    for name in ("rsaFACT-8bit", "rsaFACT-16bit", "rsaFACT-64bit"):
        h, r = host.scrape_header(cnf_text(name)), ref.scrape_header(cnf_text(name))
        for k in ("rc", "num_vars", "num_clauses", "num_bits", "product"):
            assert h[k] == r[k], (name, k)
        assert h["v1"].tolist() == r["v1"].tolist() and h["v2"].tolist() == r["v2"].tolist()


def test_parser_reports_the_widths_the_reference_drops(host, ref):
    """NDP:639 drops clause lines of width 2 or > 3 without a word; the drop-in parses the same
    clauses (parity) but counts what it ignored, and the program prints a warning."""
    text = "p cnf 5 6\n1 2 3 0\n-1 2 0\n4 0\n1 2 3 4 0\n0\n-5 -4 -3 0\n"
    got = host.parse(text)
    assert got.tolist() == [[1, 2, 3], [0, 0, 4], [-5, -4, -3]]
    assert got.tolist() == ref.parse(text).tolist()
    assert host.dropped(text) == (2, 2)
    assert host.dropped("p cnf 3 1\n1 2 3 0\n") == (0, 0)


def test_parser_differential_fuzz(host, ref, cnf_text):
    """parse_dimacs against the reference's parseDimacsString (NDP:619-642) on mutated DIMACS text: junk tokens, signs,
    overflowing integers, tabs / CR / vertical tabs, missing and doubled terminators, comment and problem lines in odd
    places.  Same clause array, bit for bit."""
    import random
    rng = random.Random(7)
    base = cnf_text("rsaFACT-16bit").splitlines()[:60]
    tokens = ["0", "-", "+5", " ", "\t", "c", "p", "x", "1.5", "2147483647", "2147483648", "-2147483648", "-2147483649",
              "99999999999999999999", "\r", "  7", "-0", "+0", "0 0", "1e3", "0x10", "--3", "+-3", "3-", "3 -", "\x0b", "\x0c"]
    for it in range(600):
        lines = list(base)
        for _ in range(rng.randint(1, 6)):
            i = rng.randrange(len(lines))
            op = rng.randrange(5)
            if op == 0:
                lines[i] = lines[i] + " " + rng.choice(tokens)
            elif op == 1:
                lines[i] = rng.choice(tokens) + " " + lines[i]
            elif op == 2:
                parts = lines[i].split(" ")
                parts.insert(rng.randrange(len(parts) + 1), rng.choice(tokens))
                lines[i] = " ".join(parts)
            elif op == 3:
                lines.insert(i, " ".join(rng.choice(tokens) for _ in range(rng.randint(0, 5))))
            else:
                lines[i] = lines[i].replace(" ", rng.choice([" ", "  ", "\t", ""]), rng.randint(1, 3))
        text = rng.choice(["\n", "\r\n"]).join(lines) + rng.choice(["", "\n", "\n\n"])
        a, b = host.parse(text), ref.parse(text)
        assert a.shape == b.shape and (a == b).all(), (it, text)


def test_header_known_answers(host, cnf_text):
    h = host.scrape_header(cnf_text("rsaFACT-16bit"))
    assert (h["rc"], h["num_vars"], h["num_clauses"], h["num_bits"], h["product"]) == (0, 336, 1296, 9, "35611")
    assert h["v1"].tolist() == [8, 7, 6, 5, 4, 3, 2, 1] and h["v2"].tolist() == list(range(16, 8, -1))
    assert host.scrape_header("p cnf 3 2\n1 2 3 0\n")["rc"] == 2          # no product line (NDP:1612-1615)
    assert host.scrape_header("c nothing\n")["rc"] == 1                    # no problem line (NDP:1602-1605)
    assert host.calculate_max_iterations(8, 5664, 1440, 17) == 4292
    assert host.calculate_max_iterations(2, 1296, 336, 9) == 969
    # header numbers std::stoi refuses: the reference prints a message and terminates (NDP:960-966, 1595-1600); the
    # drop-in reports them (rc 3; the program prints the reference's message and exits 1) -- no exception, no abort
    bad_list = cnf_text("rsaFACT-16bit").replace("[8, 7, 6,", "[8, x, 6,")
    assert host.scrape_header(bad_list)["rc"] == 3
    assert host.scrape_header("c Circuit for product = 15 [1111]\np cnf 99999999999999999999 2\n1 2 3 0\n")["rc"] == 3
    assert host.calculate_max_iterations(10, 24, 40, 1100011011) == (24 - 40 + 5 * 1100011011) % 2**32   # = 1205087743: wraps, no UB


def test_host_code_under_sanitizers(tmp_path):
    """tools/fuzz_host.cpp: parser, header scrape, CLI, naming, factor decode and the BFS JSON writer / reader fed
    mutated inputs under ASan + UBSan (a short run; `fuzz_host 30000` is the long one)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "fuzz_host")
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         "-I" + os.path.join(ROOT, "ndp-5_b200", "host"), os.path.join(ROOT, "tools", "fuzz_host.cpp"),
                         os.path.join(ROOT, "ndp-5_b200", "host", "ndp_host.cpp"), "-l:libgmp.so.10", "-o", exe],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if cc.returncode != 0 and ("asan" in cc.stdout or "ubsan" in cc.stdout):
        pytest.skip("sanitizer runtimes not installed")
    assert cc.returncode == 0, cc.stdout[-3000:]
    r = subprocess.run([exe, "400"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "iterations ok" in r.stdout, r.stdout[-3000:]


def test_naming_and_formatting_known_answers(host):
    assert host.create_problem_id("35611", 9, 4, "2025-01-01 00:00:00 UTC") == "af743d8f1e2ed8c6"
    assert host.format_filename("NDP-5_6_7", "n1234567890123.cnf", "af743d8f1e2ed8c6", "auto") == \
        "NDP-5_6_7_n12345e67890123_af743-auto.txt"
    assert host.create_bfs_filename("NDP-5_6_7", "some/dir/rsaFACT-16bit.dimacs", "af743d8f1e2ed8c6", "d60") == \
        "bfs_NDP-5_6_7_rsaFACT-16bit_af743-d60.json"
    assert host.format_duration(3725.5) == "1 hours 2 minutes 5.50 seconds\n"
    assert host.format_duration(0.004) == "0.00 seconds\n"
    assert host.format_percentage(1.5, 3.75) == "40.00%" and host.format_percentage(1, 0) == "0.00%"
    assert host.format_vector([17, 81, -88]) == "[17, 81, -88]" and host.format_vector([]) == "[]"


def test_naming_and_formatting_match_reference(host, ref):
    for args in [("35611", 9, 4, "2025-01-01 00:00:00 UTC"), ("11708736375489890129", 33, 9, "2026-10-19 23:25:23 UTC")]:
        assert host.create_problem_id(*args) == ref.create_problem_id(*args)
    for f in ["n1234567890123.cnf", "rsaFACT-16bit.dimacs", "a.b.c.cnf", "noext", "x12345.y123456789.z"]:
        for flag in ["auto", "d60", "q-5"]:
            assert host.format_filename("NDP-5_6_7", f, "502c76aff8bba126", flag) == \
                ref.format_filename("NDP-5_6_7", f, "502c76aff8bba126", flag)
            assert host.create_bfs_filename("NDP 5", "dir/" + f, "50 c7", flag) == \
                ref.create_bfs_filename("NDP 5", "dir/" + f, "50 c7", flag)
    for s in [0, 0.004, 59.999, 61, 3725.5, 90061.25, 2678400.0, 40000000.123]:
        assert host.format_duration(s) == ref.format_duration(s)
    for a, b in [(1.5, 3.75), (0, 0), (2, 3), (7, 7)]:
        assert host.format_percentage(a, b) == ref.format_percentage(a, b)
    assert host.format_vector([1, -2, 300]) == ref.format_vector([1, -2, 300])
    for w, c, v, b in [(8, 5664, 1440, 17), (2, 1296, 336, 9), (1, 10, 20, 5), (3, 100, 7, 4)]:
        assert host.calculate_max_iterations(w, c, v, b) == ref.calculate_max_iterations(w, c, v, b)


def test_cli_rules(host):
    c = host.parse_cli(["NDP-5_6_7", "in.dimacs"])
    assert (c["rc"], c["cli_flag"], c["depth"], c["max_queues"], c["override"]) == (0, "auto", 0, -1, False)
    c = host.parse_cli(["NDP-5_6_7", "in.dimacs", "-d", "60", "-s"])
    assert (c["rc"], c["cli_flag"], c["depth"], c["save"]) == (0, "d60", 60, True)
    c = host.parse_cli(["NDP-5_6_7", "in.dimacs", "-q", "1024"])
    assert (c["rc"], c["cli_flag"], c["max_queues"], c["override"]) == (0, "q1024", 1024, True)
    c = host.parse_cli(["NDP-5_6_7", "in.dimacs", "-r", "bfs_x.json", "-d", "-1"])
    assert (c["rc"], c["resume"], c["resume_file"], c["depth"], c["cli_flag"]) == (0, True, "bfs_x.json", -1, "d-1")
    c = host.parse_cli(["NDP-5_6_7", "in.dimacs", "-r", "-s"])          # -r without a file name (NDP:1498)
    assert (c["rc"], c["resume"], c["resume_file"], c["save"]) == (0, True, "", True)
    assert "cannot be used with -d" in host.parse_cli(["NDP-5_6_7", "f", "-d", "5", "-q", "7"])["error"]
    assert "cannot be used with -q" in host.parse_cli(["NDP-5_6_7", "f", "-q", "7", "-d", "5"])["error"]
    assert "Unknown or invalid argument: -z" in host.parse_cli(["NDP-5_6_7", "f", "-z"])["error"]
    assert "Unknown or invalid argument: -d" in host.parse_cli(["NDP-5_6_7", "f", "-d"])["error"]   # NDP:1516 needs a value
    assert "Missing input filename" in host.parse_cli(["NDP-5_6_7", "-d", "5"])["error"]
    assert "Usage:" in host.parse_cli(["NDP-5_6_7"])["error"]
    c = host.parse_cli(["NDP-5_6_7", "f", "-g", "8", "-x", "-q", "64"])
    assert (c["rc"], c["gpus"], c["exhaustive"], c["cli_flag"]) == (0, 8, True, "q64")


def test_factor_decode_and_gmp_check(host, golden, cnf_text):
    n = 0
    for c in golden["cases"]:
        hdr = host.scrape_header(cnf_text(c["cnf"])) if c.get("sat") else None
        for k, s in c.get("sat", {}).items():
            d1, d2, ok = host.convert(s["model"], hdr["v1"], hdr["v2"], c["N"])
            assert ok and (str(d1), str(d2)) == (s["fact1"], s["fact2"])
            bad = [x for x in s["model"] if abs(x) != hdr["v1"][-1]] + [-hdr["v1"][-1] if hdr["v1"][-1] in s["model"] else hdr["v1"][-1]]
            assert host.convert(bad, hdr["v1"], hdr["v2"], c["N"])[2] is False      # flipped lsb => "FALSE"
            n += 1
    assert n >= 10
    # a 256-bit product goes through GMP, not through machine integers
    p, q = 2 ** 127 - 1, 2 ** 127 - 1
    v1, v2 = list(range(128, 0, -1)), list(range(256, 128, -1))
    model = [v for v in v1[1:]] + [v for v in v2[1:]]        # all bits but the msb set = 2^127 - 1
    assert host.convert(model, v1, v2, str(p * q)) == (p, q, True)


def test_report_matches_reference(host, ref, tmp_path, cnf_text):
    hdr = ref.scrape_header(cnf_text("rsaFACT-16bit"))
    model = [17, 81, -88, 8, 5, 3, 1, 16, 15, 14, 12, 11, 10, 9]
    for found, bfs_s, dfs_s, fq in [(True, 1.5, 2.25, 5), (False, 0.0, 0.0, 0), (True, 3725.5, 90061.25, 0)]:
        want = ref.generate_output(found, model if found else [], bfs_s, dfs_s, 9, 336, 1296, 35611, hdr["v1"],
                                   hdr["v2"], "NDP-5_6_7", "rsaFACT-16bit.dimacs", "d60", str(tmp_path), 60, 4, 3, 4, fq)
        utc = re.search(r"Zulu time: (.*)\n", want).group(1)
        got, file_text, file_name = host.report(found, model if found else [], bfs_s, dfs_s, 9, 336, 1296, 35611,
                                                hdr["v1"], hdr["v2"], "NDP-5_6_7", "rsaFACT-16bit.dimacs", "d60", 60, 4,
                                                3, 4, fq, utc)
        head = want.split("Result saved:")[0]
        assert got == head
        saved = re.search(r"Result saved: (.*)\n", want).group(1)
        assert os.path.basename(saved) == file_name
        assert open(saved).read() == file_text


def test_report_and_naming_differential_fuzz(host, ref, tmp_path, cnf_text):
    """Randomised differential of the report block (generate_output, NDP:1127-1205) and the naming / formatting helpers
    (NDP:423-485, 533-562, 1093-1125) against the reference's own: boundary durations (59.995 s, 3599.999 s, ...), zero
    totals, 40-digit inputs, file names with and without five-digit runs, dots and directories."""
    import random
    rng = random.Random(11)
    hdr = ref.scrape_header(cnf_text("rsaFACT-16bit"))

    def seconds():
        if rng.random() < 0.3:
            return rng.choice([0.0, 59.994, 59.995, 59.999, 60.0, 3599.999, 3600.0, 86399.5])
        return rng.choice([rng.random(), rng.random() * 100, rng.random() * 10000, rng.random() * 1e6, rng.random() * 1e-4])

    names = ["rsaFACT-16bit.dimacs", "n1234567890123.cnf", "a.b.c.cnf", "noext", "dir/sub/x.dimacs", "12345.cnf", "123456.cnf",
             "x12345y678.cnf", "", ".cnf", "99999999999999999999.dimacs", "a b.cnf"]
    for it in range(120):
        found = rng.random() < 0.6
        model = [rng.choice([-1, 1]) * rng.randint(1, 336) for _ in range(rng.randint(0, 40))] if found else []
        bfs_s, dfs_s = seconds(), seconds()
        bits, nv, nc = rng.randint(1, 70), rng.randint(1, 99999), rng.randint(1, 999999)
        N = rng.choice([35611, 1, 0, 2 ** 64 - 59, 10 ** 40 + 7, 143])
        fname, flag = rng.choice(names[:8]), rng.choice(["d60", "q4096", "auto", "d-1"])
        iters, iq, wr, ws, fq = rng.randint(0, 10 ** 6), rng.randint(0, 10 ** 5), rng.randint(0, 64), rng.randint(1, 64), rng.randint(0, 10 ** 5)
        want = ref.generate_output(found, model, bfs_s, dfs_s, bits, nv, nc, N, hdr["v1"], hdr["v2"], "NDP-5_6_7", fname, flag,
                                   str(tmp_path), iters, iq, wr, ws, fq)
        utc = re.search(r"Zulu time: (.*)\n", want).group(1)
        got, file_text, file_name = host.report(found, model, bfs_s, dfs_s, bits, nv, nc, N, hdr["v1"], hdr["v2"], "NDP-5_6_7",
                                                fname, flag, iters, iq, wr, ws, fq, utc)
        assert got == want.split("Result saved:")[0], (it, fname, flag, bfs_s, dfs_s)
        saved = re.search(r"Result saved: (.*)\n", want).group(1)
        assert os.path.basename(saved) == file_name and open(saved).read() == file_text, (it, fname)
    for it in range(1000):
        x = seconds() * rng.choice([1, 1, 10, 1000])
        assert host.format_duration(x) == ref.format_duration(x), x
        a, b = seconds(), rng.choice([0.0, seconds()])
        assert host.format_percentage(a, b) == ref.format_percentage(a, b), (a, b)
    for n in names:
        for flag in ["d12", "q4096", "auto", ""]:
            for pid in ["af743d8f1e2ed8c6", "abc", ""]:
                assert host.format_filename("NDP-5_6_7", n, pid, flag) == ref.format_filename("NDP-5_6_7", n, pid, flag), (n, flag, pid)
                assert host.create_bfs_filename("NDP-5_6_7", n, pid, flag) == ref.create_bfs_filename("NDP-5_6_7", n, pid, flag)
    for n in ["35611", "0", "", "123456789012345678901234567890", "abc"]:
        for bits in [0, 9, 64]:
            for w in [1, 2, 16]:
                utc = "2025-01-01 00:00:00 UTC"
                assert host.create_problem_id(n, bits, w, utc) == ref.create_problem_id(n, bits, w, utc)
    for v in [[], [1], [1, -2, 3], list(range(-50, 50))]:
        assert host.format_vector(v) == ref.format_vector(v)


def test_report_known_answer(host, cnf_text):
    hdr = host.scrape_header(cnf_text("rsaFACT-16bit"))
    got, file_text, file_name = host.report(False, [], 1.5, 2.25, 9, 336, 1296, 35051, hdr["v1"], hdr["v2"], "NDP-5_6_7",
                                            "prime-16bit.dimacs", "auto", 969, 32, 0, 2, 0, "2026-10-19 23:25:23 UTC")
    assert got.startswith("\n\nInput Number: 35051\n              Prime!\n\n        Bits: 9\n        VARs: 336\n"
                          "     Clauses: 1296\n\n    BFS time: 1.50 seconds (40.00%)\n              1.50 seconds\n"
                          "    DFS time: 2.25 seconds (60.00%)\n              2.25 seconds\n    NDP time: 3.75 seconds\n"
                          "              3.75 seconds\n\n Total Cores: 2\n  Queue Size: 32\nProc. Queues: 32\n"
                          "       Depth: 969\n\n NDP-version: 5.6.7\n      DIMACS: prime-16bit.dimacs\n"
                          "   Zulu time: 2026-10-19 23:25:23 UTC\n  Problem ID: ")
    assert file_text == got and file_name.startswith("NDP-5_6_7_prime-16bit_") and file_name.endswith("-auto.txt")


def _items(oracle, cnf_text, name, D, ovr, Q):
    return oracle.bfs(oracle.parse(cnf_text(name)), D, ovr, Q)["frontier"]


def test_bfs_json_layout_is_stock_nlohmann_dump4(host, oracle, tmp_path, cnf_text):
    items = [(np.array([[0, 12, -20], [0, -12, 20]], np.int32), np.array([17, 81, -88], np.int32)),
             (np.zeros((0, 3), np.int32), np.zeros(0, np.int32))]
    p = str(tmp_path / "a.json")
    assert host.write_bfs_json(p, items, 1.5) == 0
    text = open(p).read()
    assert text == ('{\n    "bfs_results": [\n        {\n            "choices": [\n                17,\n                81,\n'
                    '                -88\n            ],\n            "clauses": [\n                [\n'
                    '                    0,\n                    12,\n                    -20\n                ],\n'
                    '                [\n                    0,\n                    -12,\n                    20\n'
                    '                ]\n            ]\n        },\n        {\n            "choices": [],\n'
                    '            "clauses": null\n        }\n    ],\n    "bfs_time": 1.5\n}')       # empty ClauseSet: null, as the reference writes it
    assert json.loads(text) == {"bfs_results": [{"choices": [17, 81, -88], "clauses": [[0, 12, -20], [0, -12, 20]]},
                                                {"choices": [], "clauses": None}], "bfs_time": 1.5}
    p2 = str(tmp_path / "empty.json")
    assert host.write_bfs_json(p2, [], 0.0) == 0
    assert open(p2).read() == '{\n    "bfs_time": 0.0\n}'       # empty frontier omits bfs_results
    assert host.read_bfs_json(p2) == ([], 0.0)


def test_bfs_json_round_trip_and_python_json(host, oracle, tmp_path, cnf_text):
    items = _items(oracle, cnf_text, "rsaFACT-16bit", 969, False, -1)
    p = str(tmp_path / "f.json")
    assert host.write_bfs_json(p, items, 0.0134) == 0
    back, t = host.read_bfs_json(p)
    assert t == 0.0134 and len(back) == len(items) == 32
    for (c1, h1), (c2, h2) in zip(items, back):
        assert (c1 == c2).all() and (h1 == h2).all()
    j = json.load(open(p))
    assert [e["choices"] for e in j["bfs_results"]] == [h.tolist() for _, h in items]
    assert [e["clauses"] for e in j["bfs_results"]] == [c.tolist() for c, _ in items]
    # tolerant reader: compact layout, other key order, unknown keys
    q = str(tmp_path / "g.json")
    with open(q, "w") as f:
        json.dump({"zzz": {"a": [1, {"b": None}], "s": "x\\\"y"}, "bfs_time": 2, "bfs_results":
                   [{"clauses": c.tolist(), "extra": True, "choices": h.tolist()} for c, h in items]}, f)
    back, t = host.read_bfs_json(q)
    assert t == 2.0 and all((c1 == c2).all() and (h1 == h2).all() for (c1, h1), (c2, h2) in zip(items, back))
    assert host.read_bfs_json(str(tmp_path / "missing.json")) is None
    with open(q, "w") as f:
        f.write('{"bfs_results": [{"choices": [1, 2,')
    assert host.read_bfs_json(q) is None


def test_bfs_json_cross_reads_with_reference(host, ref, oracle, tmp_path, cnf_text):
    """-s written here is readable by the reference's readBFSResults, and vice versa."""
    A = ref.parse(cnf_text("rsaFACT-16bit"))
    for D, ovr, Q in [(60, False, -1), (0, True, 32)]:
        h, info = ref.bfs_handle(A, D, ovr, Q)
        items = ref.frontier_list(h)
        # ours -> reference
        p = str(tmp_path / "ours.json")
        assert host.write_bfs_json(p, items, 0.25) == 0
        h2, t = ref.read_bfs(p)
        assert h2 and t == 0.25
        back = ref.frontier_list(h2)
        assert len(back) == len(items) and all((a == c).all() and (b == d).all() for (a, b), (c, d) in zip(items, back))
        ref.frontier_free(h2)
        # reference -> ours
        name = ref.save_bfs(h, "NDP-5_6_7", "rsaFACT-16bit.dimacs", "af743d8f1e2ed8c6", "d60", str(tmp_path), 1.5)
        assert name == host.create_bfs_filename("NDP-5_6_7", "rsaFACT-16bit.dimacs", "af743d8f1e2ed8c6", "d60")
        back, t = host.read_bfs_json(str(tmp_path / name))
        assert t == 1.5 and len(back) == len(items)
        assert all((a == c).all() and (b == d).all() for (a, b), (c, d) in zip(items, back))
        ref.frontier_free(h)


def test_bfs_json_differential_fuzz(host, ref, tmp_path):
    """Random frontiers (empty clause sets -- which the reference writes as `"clauses": null` --, empty choices, INT_MIN /
    INT_MAX literals, odd bfs_time values) through both writers and both readers: ours -> readBFSResults (NDP:600-612),
    saveBFSResults (NDP:564-598) -> ours."""
    import random
    rng = random.Random(5)
    ext = [0, 1, -1, 2 ** 31 - 1, -2 ** 31, 32766, -32767, 65536]

    def lit():
        return rng.choice(ext) if rng.random() < 0.2 else rng.randint(-5000, 5000)

    def same(a, b):
        return len(a) == len(b) and all(x[0].reshape(-1, 3).tolist() == y[0].reshape(-1, 3).tolist() and
                                        x[1].tolist() == y[1].tolist() for x, y in zip(a, b))
    for it in range(80):
        items = []
        for k in range(rng.choice([1, 2, 5, 17])):
            ncl, nch = rng.choice([0, 0, 1, 3, 40]), rng.choice([0, 1, 7, 300])
            items.append((np.array([[lit(), lit(), lit()] for _ in range(ncl)], np.int32).reshape(-1, 3),
                          np.array([lit() for _ in range(nch)], np.int32)))
        t = rng.choice([0.0, 1.5, 0.0134, 123456.789, 1e-7, 3.0, 1e20, 0.1])
        p = str(tmp_path / "ours.json")
        assert host.write_bfs_json(p, items, t) == 0
        h, t2 = ref.read_bfs(p)
        got = ref.frontier_list(h)
        ref.frontier_free(h)
        assert t2 == t and same(got, items), it
        h = ref.frontier_from_arrays(items)
        name = ref.save_bfs(h, "NDP-5_6_7", "x.cnf", "af743d8f1e2ed8c6", "q5", str(tmp_path), t)
        ref.frontier_free(h)
        back, t3 = host.read_bfs_json(str(tmp_path / name))
        assert t3 == t and same(back, items), it


def test_bfs_json_reader_rejects_fractional_literals(host, tmp_path):
    """A literal is an integer: 2.0 reads as 2 (what nlohmann's get<int>() gives), 1.5 is refused (round-1 advisor)."""
    ok = str(tmp_path / "ok.json")
    open(ok, "w").write('{"bfs_results": [{"choices": [2.0, -3], "clauses": [[1, 2.0, -3e0]]}], "bfs_time": 0.5}')
    back, t = host.read_bfs_json(ok)
    assert t == 0.5 and back[0][1].tolist() == [2, -3] and back[0][0].tolist() == [[1, 2, -3]]
    for bad in ('{"bfs_results": [{"choices": [1.5], "clauses": []}], "bfs_time": 0.5}',
                '{"bfs_results": [{"choices": [], "clauses": [[1, 2, 2.25]]}], "bfs_time": 0.5}'):
        p = str(tmp_path / "bad.json")
        open(p, "w").write(bad)
        assert host.read_bfs_json(p) is None


def test_bfs_json_writer_reports_a_full_disk(host, tmp_path):
    """A short write in mid-stream must fail the save (round-1 advisor): /dev/full accepts the open and fails every flush."""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    items = [(np.arange(1, 3001, dtype=np.int32).reshape(-1, 3), np.arange(1, 50, dtype=np.int32))] * 300   # > one 1 MB buffer
    assert host.write_bfs_json("/dev/full", items, 1.0) != 0
