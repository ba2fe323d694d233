// ndp_host.hpp -- the C++17 host side that stays a drop-in for NDP 5.6.7's own interface:
// DIMACS parser, header scrape, CLI, BFS JSON save/resume, file naming, report text and GMP
// verification of the recovered factors.  Plain host code; the hot path is reached only through
// the C ABI in include/ndp_b200.h.  Every function cites the reference lines it mirrors
// (NDP = NDP-5_6_7.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace ndph {

// ---- input ----------------------------------------------------------------------------------
// parseDimacsString, NDP:619-642: one clause per line; lines that are empty or start with 'c'/'p'
// are skipped; 1 literal -> {0,0,x}; 3 literals -> {x,y,z}; any other width is silently dropped.
// What the reference's parser does not tell its user: clause lines whose width is neither 1 nor 3
// are dropped (NDP:639), which silently changes the formula being solved.
struct DimacsDiagnostics {
    size_t dropped = 0;               // clause lines with 2 or > 3 literals
    size_t first_dropped_width = 0;
    size_t empty_lines = 0;           // lines holding only the terminating 0
};
std::vector<int32_t> parse_dimacs(const std::string& text, DimacsDiagnostics* diag = nullptr);   // n x 3 row-major

struct Header {
    bool ok_problem = false, ok_product = false;
    int num_vars = 0, num_clauses = 0, num_bits = 0;
    std::string product;            // decimal string
    std::vector<int> v1, v2;        // input variable lists, msb..lsb
    std::string error;              // a header number that std::stoi refuses: the reference's message (it then terminates, NDP:960-966)
};
// the regexes of main() (NDP:1590-1615) and ExtractInputsFromDimacs (NDP:946-990)
Header scrape_header(const std::string& text);

// calculate_max_iterations, NDP:1031-1037
int calculate_max_iterations(int world_size, int num_clauses, int num_vars, int num_bits);

// ---- CLI ------------------------------------------------------------------------------------
struct Cli {
    std::string filename;
    bool save_bfs = false;                 // -s
    bool resume = false;                   // -r [file]
    std::string resume_file;
    int max_queues = -1;                   // -q N
    int depth = 0;                         // -d N
    bool override_max_iterations = false;  // set by -q (NDP:1512)
    std::string cli_flag;                  // dN | qN | auto
    int gpus = 1;                          // -g N   (extension: devices to use; not in the reference)
    bool exhaustive = false;               // -x     (extension: no early exit; resolve every subproblem)
    std::string error;                     // non-empty => usage error (the reference std::terminate()s)
};
// parseCLIOptions, NDP:1469-1538 (same flags, same exclusivity rule, same messages)
Cli parse_cli(int argc, char** argv);
std::string usage_text(const std::string& argv0);

// ---- naming / formatting (NDP:423-485, 533-562, 1093-1125) -------------------------------------
std::string format_duration(double seconds);
std::string format_percentage(double part, double total);
std::string format_vector(const std::vector<int>& v);
std::string current_utc_time();
std::string create_problem_id(const std::string& input_number, int num_bits, int world_size, const std::string& utc);
std::string format_filename(const std::string& script, const std::string& filename, const std::string& problem_id,
                            const std::string& cli_flag);
std::string create_bfs_filename(const std::string& script, const std::string& filename, const std::string& problem_id,
                                const std::string& flag);

// ---- BFS JSON (NDP:487-612), streaming in both directions ------------------------------------
// Writer: stock nlohmann dump(4) layout, keys sorted ("bfs_results" before "bfs_time"), one array
// element per line; an empty frontier omits "bfs_results" as the reference does.  `get(k, ...)`
// supplies element k; nothing is buffered beyond one element.
using ElementGetter = std::function<bool(size_t k, const int32_t** clauses3, size_t* n_clauses,
                                         const int32_t** choices, size_t* n_choices)>;
bool write_bfs_json(const std::string& path, size_t count, const ElementGetter& get, double bfs_time,
                    std::string* err);
// Reader: any whitespace/key order; unknown keys are skipped.  Flat output arrays with offsets.
struct BfsJson {
    std::vector<int32_t> clauses3;
    std::vector<size_t> clause_off{0};
    std::vector<int32_t> choices;
    std::vector<size_t> choice_off{0};
    double bfs_time = 0.0;
    size_t count() const { return clause_off.size() - 1; }
};
bool read_bfs_json(const std::string& path, BfsJson* out, std::string* err);

// ---- factors + report ------------------------------------------------------------------------
// processVector/convert (NDP:1008-1027): bit = 1 iff the positive literal is in the model;
// product check with GMP (NDP:1153).  Returns decimal strings.
struct Factors { std::string d1, d2; bool verified = false; };
Factors decode_factors(const std::vector<int>& model, const std::vector<int>& v1, const std::vector<int>& v2,
                       const std::string& input_number);

struct Report {
    bool solution_found = false;
    std::vector<int> model;
    double bfs_seconds = 0, dfs_seconds = 0;
    int num_bits = 0, num_vars = 0, num_clauses = 0;
    std::string input_number;
    std::vector<int> v1, v2;
    std::string script_name, filename, cli_flag, output_directory;
    int iterations = 0;
    size_t initial_queue_size = 0;
    int world_rank = 0, world_size = 1, final_queue_size = 0;
    std::string utc_time;   // empty => now
    std::string node_name;  // empty => gethostname
};
// generate_output (NDP:1127-1205): returns the block printed to stdout; *file_text gets the same
// plus the " Assignments:" line; *file_name the formatFilename() result.
std::string report_text(const Report& r, std::string* file_text, std::string* file_name);

std::string read_file(const std::string& path);   // readFileToString, NDP:1456-1467
std::string working_directory();                  // getWorkingDirectory, NDP:331-344

}  // namespace ndph
