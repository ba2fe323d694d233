// ndp_main.cpp -- `NDP-5_6_7 <dimacs> [-d N | -q N] [-s] [-r [file]]`: the reference's command
// line (main, NDP-5_6_7.cpp:1540-1689) over the B200 hot path.  No MPI: the frontier is split
// statically over the GPUs of this host inside ndp_dfs_batch.  Extensions: -g N (GPUs to use,
// default 1), -x (exhaustive: no early exit).
//
// Where the reference's behaviour is a defect rather than an interface, this program differs
// on purpose: usage errors exit(1) instead of std::terminate(); normal
// completion returns 0 instead of std::terminate() (NDP:1404,1447); an empty frontier is
// reported and exits 2 instead of spinning in the dispatcher (NDP:1291); stderr is left alone
// (NDP:307-308).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <limits>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../include/ndp_b200.h"
#include "ndp_host.hpp"

using namespace ndph;

namespace {

std::string host_name() {
    char b[256];
    if (gethostname(b, sizeof b) != 0) return "localhost";
    b[sizeof b - 1] = 0;
    return b;
}

// promptForJSONFile, NDP:381-421
std::string prompt_for_json() {
    std::vector<std::string> files;
    for (const auto& e : std::filesystem::directory_iterator("."))
        if (e.is_regular_file() && e.path().string().find("bfs_") != std::string::npos)
            files.push_back(e.path().filename().string());
    if (files.empty()) {
        std::cout << "\n\n Error: No BFS results files found in the current directory.\n";
        return "";
    }
    std::cout << "\n\n Available BFS results files:\n ============================\n\n";
    for (size_t i = 0; i < files.size(); ++i) std::cout << " " << i + 1 << ": " << files[i] << "\n";
    int choice = 0;
    for (;;) {
        std::cout << "\n Enter the number of the file to resume from: ";
        if (!(std::cin >> choice) || choice < 1 || choice > (int)files.size()) {
            if (std::cin.eof()) return "";
            std::cin.clear();
            std::cin.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
            std::cout << "\n Invalid input. Please enter a number between 1 and " << files.size() << ".\n";
        } else break;
    }
    std::cout << "\n Selected BFS file: " << files[choice - 1] << "\n";
    return files[choice - 1];
}

int die(const char* what) {
    std::cout << "\n  Error: " << what << ": " << ndp_last_error() << "\n" << std::endl;
    return 1;
}

}  // namespace

int main(int argc, char** argv) {
    Cli cli = parse_cli(argc, argv);
    if (!cli.error.empty()) { std::cout << cli.error << std::endl; return 1; }
    if (cli.save_bfs) std::cout << "\n\n              saving BFS after completion!\n\n";           // NDP:1495
    const std::string script_name = std::filesystem::path(argv[0]).stem().string();                 // NDP:1566
    const std::string output_directory = working_directory();
    const std::string utc = current_utc_time();

    const std::string text = read_file(cli.filename);
    if (text.empty()) {                                                                             // NDP:1586-1589
        std::cout << "\nError: Could not open file " << cli.filename << std::endl;
        std::cout << "\n\n               Error: reading file or file is empty.\n\n\n" << std::endl;
        return 1;
    }
    const Header h = scrape_header(text);
    if (!h.error.empty()) { std::cout << h.error << std::endl; return 1; }   // the reference prints this and terminates (NDP:960-966)
    if (!h.ok_problem) { std::cout << "\nError: Could not extract number of variables and clauses from DIMACS header.\n" << std::endl; return 1; }
    if (!h.ok_product) { std::cout << "\nError: Could not extract input number from DIMACS header.\n" << std::endl; return 1; }

    int have = 0;
    if (ndp_device_count(&have) != NDP_OK || have < 1) {
        std::cout << "\n  Error: no CUDA device: " << ndp_last_error() << "\n  (the NDP hot path has no CPU fallback)\n" << std::endl;
        return 1;
    }
    const int n_gpus = std::max(1, std::min(cli.gpus, have));
    const int world_size = n_gpus + 1;   // the reference's rank 0 only dispatches (NDP:1290): np = workers + 1
    int depth = cli.depth;
    if (depth == 0 && !cli.override_max_iterations)                                                 // NDP:1606-1608
        depth = calculate_max_iterations(world_size, h.num_clauses, h.num_vars, h.num_bits);
    const std::string problem_id = create_problem_id(h.product, h.num_bits, world_size, utc);      // NDP:1617

    // banner (printWorkerCount NDP:1063-1067, printHeadNodeDetails NDP:1077-1090)
    std::cout << "\n\n NDP-version: 5.6.7\n" << std::endl;
    std::cout << " " << host_name() << " - " << n_gpus << " B200 GPU" << (n_gpus > 1 ? "s" : "") << std::endl;
    std::cout << "\n Total Cores: " << world_size << std::endl;
    std::cout << "\nInput Number: " << h.product << std::endl;
    std::cout << "        Bits: " << h.num_bits << std::endl;
    std::cout << "     Clauses: " << h.num_clauses << std::endl;
    std::cout << "        VARs: " << h.num_vars << std::endl << std::endl;
    if (cli.max_queues > 0) std::cout << "  Queue size: " << cli.max_queues << std::endl;
    else if (depth > 0) std::cout << "       Depth: " << depth << std::endl;
    std::cout << std::endl;

    DimacsDiagnostics diag;
    const std::vector<int32_t> clauses = parse_dimacs(text, &diag);                                 // NDP:1626
    if (clauses.empty()) { std::cout << "\n  Error parsing DIMACS string.\n" << std::endl; return 1; }
    if (diag.dropped)   // the reference drops these lines without a word (NDP:639); same formula, but say so
        std::cout << "     Warning: " << diag.dropped << " clause line(s) with a width other than 1 or 3 (first: "
                  << diag.first_dropped_width << " literals) are ignored, as NDP 5.6.7 does.\n" << std::endl;

    ndp_frontier* frontier = nullptr;
    int iterations = 0, task_count = 0;
    size_t initial_queue_size = 0;
    double bfs_seconds = 0.0;
    const auto bfs_start = std::chrono::high_resolution_clock::now();                               // NDP:1637
    if (cli.resume) {                                                                               // NDP:1640-1649
        std::string file = cli.resume_file.empty() ? prompt_for_json() : cli.resume_file;
        if (file.empty()) return 1;
        BfsJson j;
        std::string err;
        if (!read_bfs_json(file, &j, &err)) { std::cout << err << std::endl; return 1; }
        if (ndp_frontier_from_host(j.clauses3.data(), j.clause_off.data(), j.choices.data(), j.choice_off.data(),
                                   j.count(), &frontier) != NDP_OK)
            return die("loading the BFS results onto the GPU");
        // the DIMACS file is the root this frontier was granulated from (NDP:1584-1626): once the library has
        // checked that, the resumed run is searched exactly like a straight one
        if (ndp_frontier_attach_root(frontier, clauses.data(), clauses.size() / 3) != NDP_OK)
            std::cout << "\n        Note: " << ndp_last_error() << " -- searching the loaded clause sets as they are.\n";
        bfs_seconds = j.bfs_time;
        char b[64];
        snprintf(b, sizeof b, "%.2f", bfs_seconds);
        std::cout << "\n\n     Loading: " << file << "\n  Queue size: " << j.count() << "\n ET BFS time: " << b << " seconds\n";
    } else {
        if (ndp_bfs(clauses.data(), clauses.size() / 3, depth, cli.override_max_iterations, cli.max_queues, &frontier,
                    &iterations, &task_count, &initial_queue_size) != NDP_OK)                      // NDP:1651
            return die("BFS granulation");
        std::cout << "\r\033[K  Queue size: " << ndp_frontier_size(frontier) << " - Depth: " << iterations
                  << " - Tasks: " << task_count << std::flush;                                      // NDP:930 (final line only)
        bfs_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - bfs_start).count();
        if (cli.save_bfs) {                                                                         // NDP:1658-1662
            const std::string name = create_bfs_filename(script_name, cli.filename, problem_id, cli.cli_flag);
            bfs_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - bfs_start).count();
            std::string err;
            ElementGetter get = [&](size_t k, const int32_t** cl, size_t* ncl, const int32_t** ch, size_t* nch) {
                return ndp_frontier_get(frontier, k, cl, ncl, ch, nch) == NDP_OK;
            };
            if (!write_bfs_json(output_directory + "/" + name, ndp_frontier_size(frontier), get, bfs_seconds, &err)) {
                std::cout << "\n  Error in saving BFS results: " << err << std::endl;             // NDP:595-596
                return 1;
            }
            std::cout << "\n\n  BFS results saved to " << name << "\n";                           // NDP:593
        }
    }
    {
        char b[64];
        snprintf(b, sizeof b, "%.2f", bfs_seconds);
        std::cout << "\n\n    BFS time: " << b << " seconds  -  DFS parallel initiated.." << std::endl;   // NDP:1264
    }

    const size_t n = ndp_frontier_size(frontier);
    if (n == 0) {
        std::cout << "\n\n              BFS exhausted the search tree: the frontier is empty, nothing is left to search.\n"
                     "              (satisfied leaves are dropped during granulation, NDP-5_6_7.cpp:908,919 --\n"
                     "               use a smaller -d / -q)\n\n" << std::endl;
        ndp_frontier_free(frontier);
        return 2;
    }

    const auto dfs_start = std::chrono::high_resolution_clock::now();                               // NDP:1671
    std::vector<uint8_t> verdict(n, NDP_NOT_RUN);
    std::vector<int32_t> model((size_t)h.num_vars + clauses.size() + 16);
    size_t winner = SIZE_MAX, model_len = 0;
    ndp_stats st;
    // the reference's 1 Hz status line (NDP:1271-1288), fed by the kernels' progress word
    std::atomic<bool> dfs_done(false);
    std::thread time_printer([&]() {
        for (;;) {
            for (int i = 0; i < 10 && !dfs_done.load(); ++i) usleep(100000);
            if (dfs_done.load()) break;
            const long elapsed = (long)std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - dfs_start).count();
            size_t pulled = 0, total = 0;
            int active = 0;
            ndp_dfs_progress(&pulled, &total, &active);
            std::cout << "\033[2K\r    DFS time: " << elapsed << " seconds  -  Remaining Queue Size: " << (total - pulled)
                      << "  -  Active Workers: " << active << std::flush;
        }
    });
    const int dfs_rc = ndp_dfs_batch(frontier, n_gpus, cli.exhaustive ? 0 : 1, verdict.data(), &winner, model.data(),
                                     model.size(), &model_len, &st);
    dfs_done.store(true);
    time_printer.join();
    if (dfs_rc != NDP_OK) return die("DFS");
    const double dfs_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - dfs_start).count();

    Report r;
    r.solution_found = winner != SIZE_MAX;
    if (r.solution_found) r.model.assign(model.begin(), model.begin() + model_len);
    r.bfs_seconds = bfs_seconds;
    r.dfs_seconds = dfs_seconds;
    r.num_bits = h.num_bits; r.num_vars = h.num_vars; r.num_clauses = h.num_clauses;
    r.input_number = h.product;
    r.v1 = h.v1; r.v2 = h.v2;
    r.script_name = script_name; r.filename = cli.filename; r.cli_flag = cli.cli_flag; r.output_directory = output_directory;
    r.iterations = iterations;
    r.initial_queue_size = initial_queue_size;
    r.world_size = world_size;
    r.world_rank = r.solution_found ? 1 + (int)(winner % (size_t)n_gpus) : 0;   // the GPU that held the winner, as a worker rank
    int not_run = 0;
    for (uint8_t v : verdict) not_run += v == NDP_NOT_RUN;
    r.final_queue_size = r.solution_found ? not_run : 0;                        // subproblems never started (NDP:1300, 1402)
    std::string file_text, file_name;
    std::cout << report_text(r, &file_text, &file_name);
    const std::string full = output_directory + "/" + file_name;                                    // NDP:1200-1201
    {
        std::ofstream out(full);
        if (!out.is_open()) { std::cout << "\n  Error: Could not write to file " << full << std::endl << "\n" << std::endl; return 1; }
        out << file_text;
    }
    std::cout << "Result saved: " << full << "\n     On node: " << host_name() << "\n" << std::endl;   // NDP:1203-1204
    std::cout << "      Engine: ndp-b200, " << n_gpus << " x B200, " << st.nodes << " DFS nodes, " << st.clause_evals
              << " clause evaluations, kernel " << st.kernel_ms << " ms\n" << std::endl;
    ndp_frontier_free(frontier);
    return 0;
}
