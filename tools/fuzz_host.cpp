// tools/fuzz_host.cpp -- sanitizer fuzzing of the C++17 host drop-in (ndp-5_b200/host/ndp_host.cpp): the DIMACS
// parser, the header scrape, the streaming BFS JSON reader / writer, the CLI parser, the naming helpers and the
// factor decoder are fed mutated and random inputs under ASan + UBSan.  No GPU, no reference: the property checked is
// "never crashes, never reads out of bounds, and what the writer wrote the reader reads back".
//
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -Indp-5_b200/host \
//       tools/fuzz_host.cpp ndp-5_b200/host/ndp_host.cpp -l:libgmp.so.10 -o /tmp/fuzz_host && /tmp/fuzz_host 20000
#include "ndp_host.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <unistd.h>
#include <vector>

using namespace ndph;

static std::mt19937_64 rng(12345);
static size_t rnd(size_t n) { return n ? (size_t)(rng() % n) : 0; }

static std::string mutate(std::string s) {
    static const char* tokens[] = {"0", "-", " ", "\n", "\t", "c ", "p cnf ", "[", "]", ",", "{", "}", "\"", ":", "1e9", "-0",
                                   "2147483648", "-2147483649", "99999999999999999999", "1.5", "\\", "\"bfs_results\"",
                                   "\"clauses\"", "\"choices\"", "\"bfs_time\"", "null", "true", "\r\n", "\0x", "e", "+"};
    const int n_mut = 1 + (int)rnd(6);
    for (int m = 0; m < n_mut; ++m) {
        const size_t pos = rnd(s.size() + 1);
        switch (rnd(6)) {
        case 0: if (!s.empty()) s.erase(pos % s.size(), 1 + rnd(8)); break;
        case 1: s.insert(pos, tokens[rnd(sizeof tokens / sizeof *tokens)]); break;
        case 2: if (!s.empty()) s[pos % s.size()] = (char)rnd(256); break;
        case 3: s = s.substr(0, pos); break;                                   // truncate
        case 4: if (!s.empty()) s.insert(pos, s.substr(rnd(s.size()), rnd(64))); break;   // duplicate a chunk
        default: s.insert(pos, std::to_string((long long)rng())); break;
        }
    }
    return s;
}

static std::string sample_dimacs() {
    std::string t = "c Circuit for product = 35611 [1000101100011011]\n"
                    "c Variables for first input [msb,...,lsb]: [8, 7, 6, 5, 4, 3, 2, 1]\n"
                    "c Variables for second input [msb,...,lsb]: [16, 15, 14, 13, 12, 11, 10, 9]\n"
                    "p cnf 40 24\n";
    for (int k = 0; k < 24; ++k) {
        if (k % 5 == 4) t += std::to_string((int)rnd(40) + 1) + " 0\n";
        else t += std::to_string((rnd(2) ? 1 : -1) * ((int)rnd(40) + 1)) + " " + std::to_string((rnd(2) ? 1 : -1) * ((int)rnd(40) + 1)) +
                  " " + std::to_string((rnd(2) ? 1 : -1) * ((int)rnd(40) + 1)) + " 0\n";
    }
    return t;
}

int main(int argc, char** argv) {
    const long iters = argc > 1 ? atol(argv[1]) : 5000;
    char dir[] = "/tmp/ndp_fuzz_XXXXXX";
    if (!mkdtemp(dir)) return 2;
    const std::string path = std::string(dir) + "/f.json";
    size_t json_ok = 0, json_bad = 0, parsed = 0;
    for (long it = 0; it < iters; ++it) {
        // ---- DIMACS parser + header scrape on mutated text
        std::string text = sample_dimacs();
        if (it % 3) text = mutate(text);
        DimacsDiagnostics dg;
        std::vector<int32_t> c3 = parse_dimacs(text, &dg);
        if (c3.size() % 3) { fprintf(stderr, "parse_dimacs returned %zu ints\n", c3.size()); return 1; }
        parsed += c3.size() / 3;
        Header h = scrape_header(text);
        (void)calculate_max_iterations(1 + (int)rnd(64), h.num_clauses, h.num_vars, h.num_bits);
        // ---- factor decode with arbitrary variable lists / models
        {
            std::vector<int> model;
            for (size_t i = 0; i < rnd(80); ++i) model.push_back((int)(rng() % 200) - 100);
            Factors f = decode_factors(model, h.v1, h.v2, h.product.empty() ? "0" : h.product);
            (void)f;
        }
        // ---- naming helpers with hostile strings
        {
            std::string name = mutate("n1234567890123.cnf");
            (void)format_filename(mutate("./NDP-5_6_7"), name, create_problem_id(mutate("35611"), (int)rnd(70), (int)rnd(70), current_utc_time()),
                                  mutate("q4096"));
            (void)create_bfs_filename("NDP-5_6_7", name, "af743d8f1e2ed8c6", "d12");
            (void)format_duration((double)(rng() % 1000000) / 7.0);
            (void)format_percentage((double)rnd(100), (double)rnd(3));
        }
        // ---- CLI
        {
            std::vector<std::string> words{"NDP-5_6_7"};
            static const char* w[] = {"-d", "-q", "-s", "-r", "-g", "-x", "12", "-1", "4096", "f.cnf", "x.json", "-", "--", "99999999999", "abc", ""};
            for (size_t i = 0; i < rnd(7); ++i) words.push_back(rnd(5) ? w[rnd(sizeof w / sizeof *w)] : mutate("-q"));
            std::vector<char*> av;
            for (auto& s : words) av.push_back(const_cast<char*>(s.c_str()));
            Cli c = parse_cli((int)av.size(), av.data());
            (void)c;
        }
        // ---- JSON: write a frontier, read it back equal; then read mutated bytes
        {
            const size_t count = rnd(5);
            std::vector<std::vector<int32_t>> cl(count), ch(count);
            for (size_t k = 0; k < count; ++k) {
                cl[k].resize(3 * rnd(6));
                for (auto& x : cl[k]) x = (int32_t)(rng() % 2001) - 1000;
                ch[k].resize(rnd(9));
                for (auto& x : ch[k]) x = (int32_t)(rng() % 2001) - 1000;
            }
            std::string err;
            const double t = (double)rnd(100000) / 1000.0;
            bool ok = write_bfs_json(path, count, [&](size_t k, const int32_t** a, size_t* na, const int32_t** b, size_t* nb) {
                *a = cl[k].data(); *na = cl[k].size() / 3; *b = ch[k].data(); *nb = ch[k].size(); return true; }, t, &err);
            if (!ok) { fprintf(stderr, "write_bfs_json failed: %s\n", err.c_str()); return 1; }
            BfsJson back;
            if (!read_bfs_json(path, &back, &err)) { fprintf(stderr, "read of own output failed: %s\n", err.c_str()); return 1; }
            if (back.count() != count) { fprintf(stderr, "count %zu != %zu\n", back.count(), count); return 1; }
            for (size_t k = 0; k < count; ++k) {
                std::vector<int32_t> a(back.clauses3.begin() + 3 * back.clause_off[k], back.clauses3.begin() + 3 * back.clause_off[k + 1]);
                std::vector<int32_t> b(back.choices.begin() + back.choice_off[k], back.choices.begin() + back.choice_off[k + 1]);
                if (a != cl[k] || b != ch[k]) { fprintf(stderr, "element %zu differs after the round trip\n", k); return 1; }
            }
            std::string bytes = read_file(path);
            bytes = mutate(bytes);
            { std::ofstream o(path, std::ios::binary); o << bytes; }
            BfsJson junk;
            if (read_bfs_json(path, &junk, &err)) {
                ++json_ok;
                if (junk.clause_off.size() != junk.choice_off.size() || junk.clause_off.back() * 3 != junk.clauses3.size() ||
                    junk.choice_off.back() != junk.choices.size()) { fprintf(stderr, "inconsistent offsets accepted\n"); return 1; }
            } else ++json_bad;
        }
    }
    unlink(path.c_str());
    rmdir(dir);
    printf("fuzz_host: %ld iterations ok (%zu clauses parsed, mutated JSON accepted %zu / refused %zu)\n", iters, parsed, json_ok, json_bad);
    return 0;
}
