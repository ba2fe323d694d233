// ndp_host_capi.cpp -- C hooks over ndp_host.hpp so that the CPU test-suite (ctypes) can compare
// the host drop-in against the reference's own functions.  Built into lib/libndp_host.so.
#include <cstring>

#include "ndp_host.hpp"

using namespace ndph;

namespace {
size_t put(const std::string& s, char* out, size_t cap) {
    if (out && cap) { size_t n = s.size() < cap - 1 ? s.size() : cap - 1; memcpy(out, s.data(), n); out[n] = 0; }
    return s.size();
}
}  // namespace

extern "C" {

size_t ndph_parse_dimacs(const char* text, int32_t* out3, size_t cap_clauses) {
    std::vector<int32_t> c = parse_dimacs(text);
    if (out3 && c.size() / 3 <= cap_clauses) memcpy(out3, c.data(), c.size() * sizeof(int32_t));
    return c.size() / 3;
}
// how many clause lines parse_dimacs dropped for their width (NDP:639), and the first such width
size_t ndph_dimacs_dropped(const char* text, size_t* first_width) {
    DimacsDiagnostics d;
    parse_dimacs(text, &d);
    if (first_width) *first_width = d.first_dropped_width;
    return d.dropped;
}
int ndph_scrape_header(const char* text, int* num_vars, int* num_clauses, int* num_bits, char* product, size_t cap,
                       int32_t* v1, size_t* n1, int32_t* v2, size_t* n2, size_t vcap) {
    Header h = scrape_header(text);
    *num_vars = h.num_vars; *num_clauses = h.num_clauses; *num_bits = h.num_bits;
    put(h.product, product, cap);
    *n1 = h.v1.size(); *n2 = h.v2.size();
    for (size_t j = 0; j < h.v1.size() && j < vcap; ++j) v1[j] = h.v1[j];
    for (size_t j = 0; j < h.v2.size() && j < vcap; ++j) v2[j] = h.v2[j];
    return !h.error.empty() ? 3 : !h.ok_problem ? 1 : !h.ok_product ? 2 : 0;   // 3: a header number std::stoi refuses
}
int ndph_calculate_max_iterations(int world, int clauses, int vars, int bits) {
    return calculate_max_iterations(world, clauses, vars, bits);
}
size_t ndph_create_problem_id(const char* n, int bits, int world, const char* utc, char* out, size_t cap) {
    return put(create_problem_id(n, bits, world, utc), out, cap);
}
size_t ndph_format_filename(const char* script, const char* file, const char* pid, const char* flag, char* out, size_t cap) {
    return put(format_filename(script, file, pid, flag), out, cap);
}
size_t ndph_create_bfs_filename(const char* script, const char* file, const char* pid, const char* flag, char* out, size_t cap) {
    return put(create_bfs_filename(script, file, pid, flag), out, cap);
}
size_t ndph_format_duration(double s, char* out, size_t cap) { return put(format_duration(s), out, cap); }
size_t ndph_format_percentage(double a, double b, char* out, size_t cap) { return put(format_percentage(a, b), out, cap); }
size_t ndph_format_vector(const int32_t* v, size_t n, char* out, size_t cap) {
    return put(format_vector(std::vector<int>(v, v + n)), out, cap);
}
int ndph_convert(const int32_t* model, size_t len, const int32_t* v1, size_t n1, const int32_t* v2, size_t n2,
                 const char* input_number, char* d1, char* d2, size_t cap) {
    Factors f = decode_factors(std::vector<int>(model, model + len), std::vector<int>(v1, v1 + n1),
                               std::vector<int>(v2, v2 + n2), input_number);
    put(f.d1, d1, cap);
    put(f.d2, d2, cap);
    return f.verified ? 1 : 0;
}
// CLI: returns 0 ok / 1 usage error; fields through out-params
int ndph_parse_cli(int argc, char** argv, int* save, int* resume, char* resume_file, size_t cap, int* max_queues,
                   int* depth, int* override_mi, char* cli_flag, size_t fcap, int* gpus, int* exhaustive, char* error,
                   size_t ecap) {
    Cli c = parse_cli(argc, argv);
    *save = c.save_bfs; *resume = c.resume; *max_queues = c.max_queues; *depth = c.depth;
    *override_mi = c.override_max_iterations; *gpus = c.gpus; *exhaustive = c.exhaustive;
    put(c.resume_file, resume_file, cap);
    put(c.cli_flag, cli_flag, fcap);
    put(c.error, error, ecap);
    return c.error.empty() ? 0 : 1;
}
// JSON round trip on flat arrays
int ndph_write_bfs_json(const char* path, const int32_t* clauses3, const size_t* cl_off, const int32_t* choices,
                        const size_t* ch_off, size_t count, double bfs_time) {
    ElementGetter get = [&](size_t k, const int32_t** cl, size_t* ncl, const int32_t** ch, size_t* nch) {
        *cl = clauses3 + 3 * cl_off[k]; *ncl = cl_off[k + 1] - cl_off[k];
        *ch = choices + ch_off[k]; *nch = ch_off[k + 1] - ch_off[k];
        return true;
    };
    std::string err;
    return write_bfs_json(path, count, get, bfs_time, &err) ? 0 : 1;
}
// returns a handle; sizes via getters
void* ndph_read_bfs_json(const char* path) {
    BfsJson* j = new BfsJson;
    std::string err;
    if (!read_bfs_json(path, j, &err)) { delete j; return nullptr; }
    return j;
}
size_t ndph_json_count(void* h) { return static_cast<BfsJson*>(h)->count(); }
double ndph_json_time(void* h) { return static_cast<BfsJson*>(h)->bfs_time; }
size_t ndph_json_nclauses(void* h, size_t k) { auto* j = static_cast<BfsJson*>(h); return j->clause_off[k + 1] - j->clause_off[k]; }
size_t ndph_json_nchoices(void* h, size_t k) { auto* j = static_cast<BfsJson*>(h); return j->choice_off[k + 1] - j->choice_off[k]; }
void ndph_json_copy(void* h, size_t k, int32_t* clauses3, int32_t* choices) {
    auto* j = static_cast<BfsJson*>(h);
    memcpy(clauses3, j->clauses3.data() + 3 * j->clause_off[k], (j->clause_off[k + 1] - j->clause_off[k]) * 3 * sizeof(int32_t));
    memcpy(choices, j->choices.data() + j->choice_off[k], (j->choice_off[k + 1] - j->choice_off[k]) * sizeof(int32_t));
}
void ndph_json_free(void* h) { delete static_cast<BfsJson*>(h); }

size_t ndph_report(int solution_found, const int32_t* model, size_t len, double bfs_s, double dfs_s, int num_bits,
                   int num_vars, int num_clauses, const char* input_number, const int32_t* v1, size_t n1,
                   const int32_t* v2, size_t n2, const char* script, const char* filename, const char* cli_flag,
                   int iterations, size_t initial_queue_size, int world_rank, int world_size, int final_queue_size,
                   const char* utc, char* out, size_t cap, char* file_text, size_t fcap, char* file_name, size_t ncap) {
    Report r;
    r.solution_found = solution_found != 0;
    r.model.assign(model, model + len);
    r.bfs_seconds = bfs_s; r.dfs_seconds = dfs_s;
    r.num_bits = num_bits; r.num_vars = num_vars; r.num_clauses = num_clauses;
    r.input_number = input_number;
    r.v1.assign(v1, v1 + n1); r.v2.assign(v2, v2 + n2);
    r.script_name = script; r.filename = filename; r.cli_flag = cli_flag;
    r.iterations = iterations; r.initial_queue_size = initial_queue_size;
    r.world_rank = world_rank; r.world_size = world_size; r.final_queue_size = final_queue_size;
    r.utc_time = utc ? utc : "";
    std::string ft, fn;
    std::string s = report_text(r, &ft, &fn);
    put(ft, file_text, fcap);
    put(fn, file_name, ncap);
    return put(s, out, cap);
}
}  // extern "C"
