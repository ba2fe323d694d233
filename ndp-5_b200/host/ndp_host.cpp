// ndp_host.cpp -- see ndp_host.hpp.  NDP = the reference's NDP-5_6_7.cpp.
#include "ndp_host.hpp"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <regex>
#include <sstream>
#include <unordered_set>

#include <unistd.h>

// The image ships the GMP runtime (libgmp.so.10) without headers: declare the handful of entry
// points used for the product check (NDP:997-1027, 1153).  Link with -l:libgmp.so.10.
extern "C" {
typedef struct { int alloc; int size; unsigned long* d; } ndph_mpz;
void __gmpz_init(ndph_mpz*);
void __gmpz_clear(ndph_mpz*);
int __gmpz_set_str(ndph_mpz*, const char*, int);
char* __gmpz_get_str(char*, int, const ndph_mpz*);
void __gmpz_mul(ndph_mpz*, const ndph_mpz*, const ndph_mpz*);
void __gmpz_mul_2exp(ndph_mpz*, const ndph_mpz*, unsigned long);
void __gmpz_add_ui(ndph_mpz*, const ndph_mpz*, unsigned long);
int __gmpz_cmp(const ndph_mpz*, const ndph_mpz*);
}

namespace ndph {

// ------------------------------------------------------------------------------------------ input
namespace {
// one `iss >> int` of the reference's parser: skips blanks, optional sign, digits; fails (and
// ends the clause) on anything else or on int overflow
bool next_int(const char*& p, const char* end, int& out) {
    while (p < end && std::isspace((unsigned char)*p)) ++p;
    const char* q = p;
    if (q < end && (*q == '+' || *q == '-')) ++q;
    if (q >= end || !std::isdigit((unsigned char)*q)) return false;
    long long v = 0;
    bool big = false;
    while (q < end && std::isdigit((unsigned char)*q)) {
        if (!big) { v = v * 10 + (*q - '0'); big = v > (long long)INT_MAX + 1; }
        ++q;
    }
    if (*p == '-') v = -v;
    if (big || v > INT_MAX || v < INT_MIN) return false;
    out = (int)v;
    p = q;
    return true;
}
}  // namespace

std::vector<int32_t> parse_dimacs(const std::string& text, DimacsDiagnostics* diag) {
    std::vector<int32_t> out;
    if (diag) *diag = DimacsDiagnostics();
    const char* p = text.data();
    const char* end = p + text.size();
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        if (eol > p && *p != 'c' && *p != 'p') {             // NDP:625
            int lits[3] = {0, 0, 0};
            size_t n = 0;
            const char* q = p;
            int v;
            while (next_int(q, eol, v) && v != 0) {          // NDP:630
                if (n < 3) lits[n] = v;
                ++n;
            }
            if (n == 1) { out.push_back(0); out.push_back(0); out.push_back(lits[0]); }        // NDP:632-634
            else if (n == 3) { out.push_back(lits[0]); out.push_back(lits[1]); out.push_back(lits[2]); }  // NDP:635-637
            else if (diag) {                                                                    // NDP:639: dropped without a word
                if (n == 0) ++diag->empty_lines;
                else {
                    if (!diag->dropped) diag->first_dropped_width = n;
                    ++diag->dropped;
                }
            }
        }
        p = eol + 1;
    }
    return out;
}

Header scrape_header(const std::string& text) {
    Header h;
    std::smatch m;
    static const std::regex re_problem(R"(p cnf ([0-9]+) ([0-9]+))");                                   // NDP:1592
    static const std::regex re_product(R"(Circuit for product = ([0-9]+) \[)");                         // NDP:1591
    static const std::regex re_bits(R"(Variables for second input \[msb,...,lsb\]: \[.*?,\s*(\d+)\])");  // NDP:1598
    static const std::regex re_first(R"(Variables for first input \[msb,...,lsb\]: \[(.*?)\])");        // NDP:948
    static const std::regex re_second(R"(Variables for second input \[msb,...,lsb\]: \[(.*?)\])");      // NDP:949
    // std::stoi as the reference uses it (NDP:957-959, 1595-1600), but a number that does not convert is reported in
    // Header::error with the reference's own message instead of leaving the function as an exception
    auto to_int = [&](const std::string& tok, int& dst, const char* lead) {
        try { dst = std::stoi(tok); return true; }
        catch (const std::invalid_argument&) { if (h.error.empty()) h.error = std::string(lead) + "Invalid argument while converting to int: " + tok; }
        catch (const std::out_of_range&) { if (h.error.empty()) h.error = std::string(lead) + "Out of range error while converting to int: " + tok; }
        return false;
    };
    if (std::regex_search(text, m, re_problem)) {
        h.ok_problem = to_int(m[1].str(), h.num_vars, "\n") & to_int(m[2].str(), h.num_clauses, "\n");
        if (h.ok_problem && std::regex_search(text, m, re_bits)) to_int(m[1].str(), h.num_bits, "\n");
    }
    if (std::regex_search(text, m, re_product)) { h.ok_product = true; h.product = m[1].str(); }
    auto list = [&](const std::regex& re, std::vector<int>& dst, const char* lead) {
        if (!std::regex_search(text, m, re)) return;
        std::istringstream iss(m[1].str());
        std::string tok;
        while (std::getline(iss, tok, ',')) {                                  // NDP:957-959
            int v = 0;
            if (!to_int(tok, v, lead)) return;                                 // (the reference rethrows here: NDP:960-966)
            dst.push_back(v);
        }
    };
    list(re_first, h.v1, "\n");     // NDP:961 prints a leading newline for the first list only
    list(re_second, h.v2, "");
    return h;
}

int calculate_max_iterations(int world_size, int num_clauses, int num_vars, int num_bits) {
    // NDP:1034: num_clauses - num_vars + (world_size / 2) * num_bits, in unsigned arithmetic so that a nonsense header
    // (a 10-digit "bit count") wraps as the reference's compiled code does instead of being undefined behaviour
    return (int)((unsigned)num_clauses - (unsigned)num_vars + (unsigned)(world_size / 2) * (unsigned)num_bits);
}

std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return "";
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string working_directory() {
    char buf[PATH_MAX];
    return getcwd(buf, sizeof buf) ? std::string(buf) : std::string();
}

// -------------------------------------------------------------------------------------------- CLI
std::string usage_text(const std::string& argv0) {
    return "\n               Usage: " + argv0 +
           "               <filename> [-s spot_instance_ready] [-r resume_from_bfs] [-d depth | -q max_queues]\n"
           "\n"
           "               Note: -d (depth) and -q (max_queues) cannot be used together.\n\n\n";
}

Cli parse_cli(int argc, char** argv) {
    Cli c;
    if (argc < 2) { c.error = usage_text(argc ? argv[0] : "NDP"); return c; }                 // NDP:1472-1479
    c.filename = argv[1];
    if (!c.filename.empty() && c.filename[0] == '-') {                                        // NDP:1484-1488
        c.error = "\n\n               [ERROR] Missing input filename.\n"
                  "               Please provide a valid filename as the first argument.\n\n\n";
        return c;
    }
    bool has_depth = false, has_queues = false;
    auto to_int = [&](const char* s, int& dst) {
        try { dst = std::stoi(s); return true; } catch (...) { c.error = std::string("\n\n               [ERROR] Not a number: ") + s + "\n"; return false; }
    };
    for (int i = 2; i < argc; ++i) {
        const std::string opt = argv[i];
        if (opt == "-s") {
            c.save_bfs = true;                                                                // NDP:1493-1495
        } else if (opt == "-r") {
            c.resume = true;                                                                  // NDP:1496-1504
            if (i + 1 < argc && argv[i + 1][0] != '-') c.resume_file = argv[++i];
        } else if (opt == "-q" && i + 1 < argc) {                                             // NDP:1506-1514
            if (has_depth) { c.error = "\n\n               Error: Invalid combination: -q (max_queues) cannot be used with -d (depth).\n\n\n"; return c; }
            if (!to_int(argv[++i], c.max_queues)) return c;
            c.override_max_iterations = true;
            c.cli_flag += "q" + std::to_string(c.max_queues);
            has_queues = true;
        } else if (opt == "-d" && i + 1 < argc) {                                             // NDP:1516-1523
            if (has_queues) { c.error = "\n\n               Error: Invalid combination: -d (depth) cannot be used with -q (max_queues).\n\n\n"; return c; }
            if (!to_int(argv[++i], c.depth)) return c;
            c.cli_flag += "d" + std::to_string(c.depth);
            has_depth = true;
        } else if (opt == "-g" && i + 1 < argc) {                                             // extension
            if (!to_int(argv[++i], c.gpus)) return c;
        } else if (opt == "-x") {                                                             // extension
            c.exhaustive = true;
        } else {                                                                              // NDP:1525-1533
            c.error = "\n\n               [ERROR] Unknown or invalid argument: " + opt + "\n" +
                      "\n               Usage:\n"
                      "               <filename> [-s spot_instance_ready] [-r resume_from_bfs] [-d depth | -q max_queues]\n"
                      "\n"
                      "               Note: -d (depth) and -q (max_queues) cannot be used together.\n\n\n";
            return c;
        }
    }
    if (!has_depth && !has_queues) c.cli_flag = "auto";                                       // NDP:1535-1537
    return c;
}

// ------------------------------------------------------------------------------- naming / format
static std::string fixed2(double v) {
    char b[64];
    snprintf(b, sizeof b, "%.2f", v);
    return b;
}

std::string format_duration(double seconds) {                                                 // NDP:423-442
    const int months = (int)(seconds / (60 * 60 * 24 * 30));
    seconds -= months * 60 * 60 * 24 * 30;
    const int days = (int)(seconds / (60 * 60 * 24));
    seconds -= days * 60 * 60 * 24;
    const int hours = (int)(seconds / (60 * 60));
    seconds -= hours * 60 * 60;
    const int minutes = (int)(seconds / 60);
    seconds -= minutes * 60;
    std::string s;
    if (months > 0) s += std::to_string(months) + " months ";
    if (days > 0) s += std::to_string(days) + " days ";
    if (hours > 0) s += std::to_string(hours) + " hours ";
    if (minutes > 0) s += std::to_string(minutes) + " minutes ";
    return s + fixed2(seconds) + " seconds\n";
}

std::string format_percentage(double part, double total) {                                    // NDP:1093-1099
    return fixed2(total > 0.0 ? part / total * 100.0 : 0.0) + "%";
}

std::string format_vector(const std::vector<int>& v) {                                        // NDP:1114-1125
    std::string s = "[";
    for (size_t i = 0; i < v.size(); ++i) {
        s += std::to_string(v[i]);
        if (i + 1 < v.size()) s += ", ";
    }
    return s + "]";
}

std::string current_utc_time() {                                                              // NDP:444-453
    std::time_t t = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
    std::tm tm = *std::gmtime(&t);
    char b[64];
    strftime(b, sizeof b, "%Y-%m-%d %H:%M:%S UTC", &tm);
    return b;
}

std::string create_problem_id(const std::string& input_number, int num_bits, int world_size, const std::string& utc) {
    const std::string data = input_number + "-" + std::to_string(num_bits) + "-" + std::to_string(world_size) + "-" + utc;
    char b[32];
    snprintf(b, sizeof b, "%zx", std::hash<std::string>{}(data));                             // NDP:461-465
    return std::string(b).substr(0, 16);
}

static std::string strip_extension(std::string s) {
    const size_t pos = s.find_last_of('.');
    return pos == std::string::npos ? s : s.substr(0, pos);
}

std::string format_filename(const std::string& script, const std::string& filename, const std::string& problem_id,
                            const std::string& cli_flag) {                                    // NDP:468-485
    static const std::regex five_then_more(R"((\d{5})(\d+))");
    const std::string stem = std::regex_replace(strip_extension(filename), five_then_more, "$1e$2");
    return script + "_" + stem + "_" + problem_id.substr(0, 5) + "-" + cli_flag + ".txt";
}

std::string create_bfs_filename(const std::string& script, const std::string& filename, const std::string& problem_id,
                                const std::string& flag) {                                    // NDP:533-562
    auto sanitize = [](std::string s) {
        for (char& ch : s)
            if (!std::isalnum((unsigned char)ch) && ch != '_' && ch != '-') ch = '_';
        return s;
    };
    const std::string base = sanitize(strip_extension(std::filesystem::path(filename).filename().string()));
    return "bfs_" + script + "_" + base + "_" + sanitize(problem_id.substr(0, 5)) + "-" + sanitize(flag) + ".json";
}

// ------------------------------------------------------------------------------------------ JSON
namespace {
struct Out {   // buffered text sink
    FILE* f;
    std::string buf;
    explicit Out(FILE* f_) : f(f_) { buf.reserve(1 << 20); }
    bool failed = false;   // latched: a short write anywhere in the stream (ENOSPC, quota) fails the whole save
    void put(const char* s) { buf += s; if (buf.size() > (1u << 20) - 256) flush(); }
    void put_int(int v) { char b[16]; auto r = std::to_chars(b, b + sizeof b, v); buf.append(b, r.ptr); if (buf.size() > (1u << 20) - 256) flush(); }
    bool flush() {
        size_t n = fwrite(buf.data(), 1, buf.size(), f);
        if (n != buf.size() || ferror(f)) failed = true;
        buf.clear();
        return !failed;
    }
};
std::string json_double(double v) {   // shortest round-trip form, always with a fraction or exponent
    char b[64];
    auto r = std::to_chars(b, b + sizeof b, v);
    std::string s(b, r.ptr);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}
}  // namespace

bool write_bfs_json(const std::string& path, size_t count, const ElementGetter& get, double bfs_time,
                    std::string* err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { if (err) *err = "Could not open file for writing: " + path; return false; }
    Out o(f);
    o.put("{\n");
    if (count) {   // serializeBFSResults only creates the key when it pushes an entry (NDP:492-506)
        o.put("    \"bfs_results\": [\n");
        for (size_t k = 0; k < count; ++k) {
            const int32_t *cl = nullptr, *ch = nullptr;
            size_t ncl = 0, nch = 0;
            if (!get(k, &cl, &ncl, &ch, &nch)) { fclose(f); if (err) *err = "frontier element unavailable"; return false; }
            o.put("        {\n            \"choices\": [");
            if (nch) {
                o.put("\n");
                for (size_t j = 0; j < nch; ++j) { o.put("                "); o.put_int(ch[j]); o.put(j + 1 < nch ? ",\n" : "\n"); }
                o.put("            ],\n");
            } else o.put("],\n");
            // an empty ClauseSet leaves serialized_clauses a default-constructed json: stock nlohmann prints null (NDP:498-503)
            o.put(ncl ? "            \"clauses\": [" : "            \"clauses\": null\n");
            if (ncl) {
                o.put("\n");
                for (size_t c = 0; c < ncl; ++c) {
                    o.put("                [\n");
                    for (int j = 0; j < 3; ++j) { o.put("                    "); o.put_int(cl[3 * c + j]); o.put(j < 2 ? ",\n" : "\n"); }
                    o.put(c + 1 < ncl ? "                ],\n" : "                ]\n");
                }
                o.put("            ]\n");
            }
            o.put(k + 1 < count ? "        },\n" : "        }\n");
        }
        o.put("    ],\n");
    }
    o.put("    \"bfs_time\": ");
    o.put(json_double(bfs_time).c_str());
    o.put("\n}");
    bool ok = o.flush();
    ok = fclose(f) == 0 && ok;
    if (!ok && err) *err = "write failed: " + path;
    return ok;
}

namespace {
struct In {   // buffered character source with one-character look-ahead
    FILE* f;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    explicit In(FILE* f_) : f(f_), buf(1 << 20) {}
    int peek() {
        if (pos == len) { len = fread(buf.data(), 1, buf.size(), f); pos = 0; if (!len) return EOF; }
        return (unsigned char)buf[pos];
    }
    int get() { int c = peek(); if (c != EOF) ++pos; return c; }
    void ws() { while (std::isspace(peek())) ++pos; }
    bool expect(char c) { ws(); if (peek() != c) return false; ++pos; return true; }
    bool string(std::string& s) {
        ws();
        if (get() != '"') return false;
        s.clear();
        for (int c; (c = get()) != EOF;) {
            if (c == '"') return true;
            if (c == '\\') { int d = get(); if (d == EOF) return false; s += (char)d; }
            else s += (char)c;
        }
        return false;
    }
    bool number(double& v, bool& integral, long long& iv) {
        ws();
        std::string t;
        for (int c = peek(); c != EOF && (std::isdigit(c) || c == '-' || c == '+' || c == '.' || c == 'e' || c == 'E'); c = peek()) { t += (char)c; ++pos; }
        if (t.empty()) return false;
        integral = t.find_first_of(".eE") == std::string::npos;
        char* e = nullptr;
        if (integral) { iv = strtoll(t.c_str(), &e, 10); v = (double)iv; }
        else { v = strtod(t.c_str(), &e); iv = (long long)v; }
        return e && *e == 0;
    }
    bool skip_value() {   // any JSON value
        ws();
        int c = peek();
        if (c == '"') { std::string s; return string(s); }
        if (c == '{' || c == '[') {
            const char close = c == '{' ? '}' : ']';
            ++pos;
            ws();
            if (peek() == close) { ++pos; return true; }
            for (;;) {
                if (close == '}') { std::string k; if (!string(k) || !expect(':')) return false; }
                if (!skip_value()) return false;
                ws();
                int d = get();
                if (d == close) return true;
                if (d != ',') return false;
            }
        }
        if (std::isalpha(c)) { while (std::isalpha(peek())) ++pos; return true; }   // true/false/null
        double v; bool i; long long iv;
        return number(v, i, iv);
    }
    template <class F> bool array(F&& element) {   // '[' element (',' element)* ']'
        if (!expect('[')) return false;
        ws();
        if (peek() == ']') { ++pos; return true; }
        for (;;) {
            if (!element()) return false;
            ws();
            int d = get();
            if (d == ']') return true;
            if (d != ',') return false;
        }
    }
    template <class F> bool object(F&& member) {   // member(key) consumes the value
        if (!expect('{')) return false;
        ws();
        if (peek() == '}') { ++pos; return true; }
        for (;;) {
            std::string k;
            if (!string(k) || !expect(':') || !member(k)) return false;
            ws();
            int d = get();
            if (d == '}') return true;
            if (d != ',') return false;
        }
    }
    bool integer(int32_t& out) {
        double v; bool i; long long iv;
        if (!number(v, i, iv) || (!i && v != (double)iv) || iv > INT_MAX || iv < INT_MIN) return false;   // 2.0 is 2; 1.5 is not a literal
        out = (int32_t)iv;
        return true;
    }
};
}  // namespace

bool read_bfs_json(const std::string& path, BfsJson* out, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { if (err) *err = "\n    Error: Could not open " + path + " for reading.\n"; return false; }   // NDP:602-605
    In in(f);
    *out = BfsJson();
    bool ok = in.object([&](const std::string& key) {
        if (key == "bfs_time") { bool i; long long iv; return in.number(out->bfs_time, i, iv); }             // NDP:529
        if (key != "bfs_results") return in.skip_value();
        return in.array([&]() {                                                                             // NDP:516-527
            size_t n_before_cl = out->clauses3.size(), n_before_ch = out->choices.size();
            bool entry_ok = in.object([&](const std::string& k2) {
                if (k2 == "choices") return in.array([&]() { int32_t v; if (!in.integer(v)) return false; out->choices.push_back(v); return true; });
                if (k2 == "clauses") {
                    // what the reference writes for an empty ClauseSet; its reader iterates a null as empty (NDP:519-522)
                    in.ws();
                    if (in.peek() == 'n') return in.skip_value();
                    return in.array([&]() {
                        int got = 0;
                        bool a = in.array([&]() { int32_t v; if (!in.integer(v)) return false; if (got < 3) out->clauses3.push_back(v); ++got; return true; });
                        return a && got == 3;
                    });
                }
                return in.skip_value();
            });
            (void)n_before_cl; (void)n_before_ch;
            out->clause_off.push_back(out->clauses3.size() / 3);
            out->choice_off.push_back(out->choices.size());
            return entry_ok;
        });
    });
    fclose(f);
    if (!ok && err) *err = "malformed BFS results file: " + path;
    return ok;
}

// -------------------------------------------------------------------------------- factors / report
namespace {
std::string bits_to_decimal(const std::string& bits) {   // binaryStringToDecimal, NDP:997-1006
    ndph_mpz z;
    __gmpz_init(&z);
    for (char c : bits) {
        __gmpz_mul_2exp(&z, &z, 1);
        if (c == '1') __gmpz_add_ui(&z, &z, 1);
    }
    char* s = __gmpz_get_str(nullptr, 10, &z);
    std::string r(s);
    free(s);
    __gmpz_clear(&z);
    return r;
}
}  // namespace

Factors decode_factors(const std::vector<int>& model, const std::vector<int>& v1, const std::vector<int>& v2,
                       const std::string& input_number) {
    const std::unordered_set<int> set(model.begin(), model.end());                            // NDP:1011
    auto word = [&](const std::vector<int>& vars) {
        std::string bits;
        for (int k : vars) bits += set.count(k) ? '1' : '0';                                  // NDP:1012-1013
        return bits_to_decimal(bits);
    };
    Factors fz;
    fz.d1 = word(v1);
    fz.d2 = word(v2);
    ndph_mpz a, b, n, p;
    __gmpz_init(&a); __gmpz_init(&b); __gmpz_init(&n); __gmpz_init(&p);
    __gmpz_set_str(&a, fz.d1.c_str(), 10);
    __gmpz_set_str(&b, fz.d2.c_str(), 10);
    __gmpz_set_str(&n, input_number.c_str(), 10);
    __gmpz_mul(&p, &a, &b);
    fz.verified = __gmpz_cmp(&p, &n) == 0;                                                    // NDP:1153
    __gmpz_clear(&a); __gmpz_clear(&b); __gmpz_clear(&n); __gmpz_clear(&p);
    return fz;
}

std::string report_text(const Report& r, std::string* file_text, std::string* file_name) {
    const double ndp = r.bfs_seconds + r.dfs_seconds;
    const int processed = (int)r.initial_queue_size - r.final_queue_size;                     // NDP:1142
    std::string s;
    if (r.solution_found) {                                                                   // NDP:1144-1154
        const Factors fz = decode_factors(r.model, r.v1, r.v2, r.input_number);
        s += "\n\n                                  Process " + std::to_string(r.world_rank) + " found a solution!\n";
        s += "                    Remaining Queue Size: " + std::to_string(r.final_queue_size) + "\n";
        s += "\nInput Number: " + r.input_number + "\n";
        s += "      FACT 1: " + fz.d1 + "\n";
        s += "      FACT 2: " + fz.d2 + "\n";
        s += fz.verified ? "              verified.\n" : "              FALSE\n";
    } else {                                                                                  // NDP:1155-1158
        s += "\n\nInput Number: " + r.input_number + "\n";
        s += "              Prime!\n";
    }
    s += "\n        Bits: " + std::to_string(r.num_bits);                                     // NDP:1160-1162
    s += "\n        VARs: " + std::to_string(r.num_vars);
    s += "\n     Clauses: " + std::to_string(r.num_clauses);
    s += "\n\n    BFS time: " + fixed2(r.bfs_seconds) + " seconds (" + format_percentage(r.bfs_seconds, ndp) + ")\n";
    s += "              " + format_duration(r.bfs_seconds);                                   // NDP:1164-1166
    s += "    DFS time: " + fixed2(r.dfs_seconds) + " seconds (" + format_percentage(r.dfs_seconds, ndp) + ")\n";
    s += "              " + format_duration(r.dfs_seconds);                                   // NDP:1168-1170
    s += "    NDP time: " + fixed2(ndp) + " seconds\n";
    s += "              " + format_duration(ndp);                                             // NDP:1172-1173
    s += "\n Total Cores: " + std::to_string(r.world_size) + "\n";                            // NDP:1175-1178
    s += "  Queue Size: " + std::to_string(r.initial_queue_size) + "\n";
    s += "Proc. Queues: " + std::to_string(processed) + "\n";
    s += "       Depth: " + std::to_string(r.iterations) + "\n";
    s += "\n NDP-version: 5.6.7\n";                                                           // NDP:315, 1180
    s += "      DIMACS: " + r.filename + "\n";
    const std::string utc = r.utc_time.empty() ? current_utc_time() : r.utc_time;
    s += "   Zulu time: " + utc + "\n";                                                       // NDP:1183-1186
    const std::string pid = create_problem_id(r.input_number, r.num_bits, r.world_size, utc);
    s += "  Problem ID: " + pid + "\n";
    if (file_text) {
        *file_text = s;
        if (r.solution_found) *file_text += "\n Assignments: " + format_vector(r.model) + "\n";   // NDP:1194-1196
    }
    if (file_name)
        *file_name = format_filename(r.script_name, std::filesystem::path(r.filename).filename().string(), pid,
                                     r.cli_flag);                                             // NDP:1198-1199
    return s;
}

}  // namespace ndph
