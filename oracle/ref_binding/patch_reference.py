#!/usr/bin/env python3
"""Apply INTEGRATION.md §1-§2 to the reference translation unit -- the reference-side binding a maintainer would
add -- so that it can be COMPILED and RUN against libndp_b200.so ("the reference-side binding is
never built").  Test infrastructure: the output is written where the caller says (tests: tmp_path; the recipe in
oracle/Makefile: oracle/_ref/), never into the repo's sources, and /root/reference itself is not touched.

    python oracle/ref_binding/patch_reference.py /root/reference/NDP-5_6_7.cpp out/NDP-5_6_7_b200.cpp

Three textual replacements, each anchored on the reference's own line (the script fails if an anchor is missing
or ambiguous, i.e. if the reference changed):
  1. after `#include <mpi.h>`            : + #include "ndp_b200.h" and one file-scope frontier handle
  2. the call at NDP-5_6_7.cpp:1651      : Satisfy_iterative_BFS(...)  ->  ndp_bfs(...)        
  3. the call at NDP-5_6_7.cpp:1673-1677 : process_queue(...)          ->  ndp_dfs_batch(...) + generate_output(...)
                                                                                              
Everything else -- CLI, header scrape, DIMACS parser, -s/-r JSON, report, GMP check -- is the reference's own code.
"""
import re
import sys

INCLUDE_ANCHOR = "#include <mpi.h>"
INCLUDE_NEW = '''#include <mpi.h>
#include "ndp_b200.h"   // [ndp_b200 binding] the B200 hot path behind a C ABI
static ndp_frontier* g_ndp_frontier = nullptr;   // [ndp_b200 binding] the BFS frontier, device resident
'''

BFS_ANCHOR = ("auto [results, task_count_] = Satisfy_iterative_BFS(clauses, depth, override_max_iterations, iterations, "
              "max_queues, world_rank, initial_queue_size);")
BFS_NEW = '''// [ndp_b200 binding] was: Satisfy_iterative_BFS(clauses, depth, override_max_iterations, iterations, max_queues, world_rank, initial_queue_size)
		std::queue<std::pair<ClauseSet, std::vector<int>>> results;
		int task_count_ = 0;
		{
			std::vector<int> flat = flattenClauseSet(clauses);
			if (ndp_bfs(flat.data(), clauses.size(), depth, override_max_iterations, max_queues, &g_ndp_frontier,
			            &iterations, &task_count_, &initial_queue_size) != NDP_OK)
				throw std::runtime_error(ndp_last_error());
			if (spot_instance_ready) {   // -s: saveBFSResults walks the queue (NDP:564), so materialise it
				for (size_t k = 0; k < ndp_frontier_size(g_ndp_frontier); ++k) {
					const int32_t *cl, *ch; size_t ncl, nch;
					if (ndp_frontier_get(g_ndp_frontier, k, &cl, &ncl, &ch, &nch) != NDP_OK)
						throw std::runtime_error(ndp_last_error());
					results.emplace(unflattenClauseSet(std::vector<int>(cl, cl + 3 * ncl)), std::vector<int>(ch, ch + nch));
				}
			}
		}'''

DFS_ANCHOR_RE = re.compile(r"process_queue\(bfs_queue, true, input_number, num_bits, num_vars, num_clauses,\s*"
                           r"v1, v2, bfs_start, dfs_start, task_count, script_name, filename,\s*"
                           r"cli_flag, output_directory, iterations,\s*"
                           r"initial_queue_size, world_rank, world_size, MPI_COMM_WORLD,\s*"
                           r"bfs_duration_chrono\);")
DFS_NEW = '''// [ndp_b200 binding] was: process_queue(bfs_queue, true, ...) -- dispatcher + workers (NDP:1229-1454)
	{
		if (!g_ndp_frontier) {   // -r: the queue came from readBFSResults; hand it to the library with its root
			std::vector<int> cl3, ch;
			std::vector<size_t> cl_off(1, 0), ch_off(1, 0);
			auto q = bfs_queue;
			while (!q.empty()) {
				std::vector<int> flat = flattenClauseSet(q.front().first);
				cl3.insert(cl3.end(), flat.begin(), flat.end());
				ch.insert(ch.end(), q.front().second.begin(), q.front().second.end());
				cl_off.push_back(cl3.size() / 3);
				ch_off.push_back(ch.size());
				q.pop();
			}
			if (ndp_frontier_from_host(cl3.data(), cl_off.data(), ch.data(), ch_off.data(), cl_off.size() - 1,
			                           &g_ndp_frontier) != NDP_OK)
				throw std::runtime_error(ndp_last_error());
			std::vector<int> root = flattenClauseSet(clauses);
			ndp_frontier_attach_root(g_ndp_frontier, root.data(), clauses.size());   // refused => searched as loaded
			if (initial_queue_size == 0) initial_queue_size = ndp_frontier_size(g_ndp_frontier);
		}
		int n_gpus = 1;
		if (const char* e = getenv("NDP_GPUS")) n_gpus = std::max(1, atoi(e));
		const size_t n_front = ndp_frontier_size(g_ndp_frontier);
		std::vector<uint8_t> verdict(n_front ? n_front : 1, NDP_NOT_RUN);
		std::vector<int32_t> model((size_t)num_vars + 16);
		size_t winner = SIZE_MAX, model_len = 0;
		ndp_stats st;
		if (ndp_dfs_batch(g_ndp_frontier, n_gpus, /*early_exit=*/1, verdict.data(), &winner, model.data(), model.size(),
		                  &model_len, &st) != NDP_OK)
			throw std::runtime_error(ndp_last_error());
		const bool found = winner != SIZE_MAX;
		std::vector<std::vector<int>> sols;
		if (found) sols.emplace_back(model.begin(), model.begin() + model_len);   // = final_choices_i (NDP:1430-1431)
		generate_output(found, sols, bfs_duration_chrono, std::chrono::high_resolution_clock::now() - dfs_start,
		                num_bits, num_vars, num_clauses, input_number, v1, v2, script_name, filename, cli_flag,
		                output_directory, iterations, initial_queue_size, world_rank, world_size,
		                found ? int(n_front - winner - 1) : 0);
		ndp_frontier_free(g_ndp_frontier);
		g_ndp_frontier = nullptr;
	}'''


def patch(src):
    assert src.count(INCLUDE_ANCHOR) == 1, "anchor 1 (#include <mpi.h>) not found exactly once"
    src = src.replace(INCLUDE_ANCHOR, INCLUDE_NEW, 1)
    assert src.count(BFS_ANCHOR) == 1, "anchor 2 (the Satisfy_iterative_BFS call, NDP:1651) not found exactly once"
    src = src.replace(BFS_ANCHOR, BFS_NEW, 1)
    hits = DFS_ANCHOR_RE.findall(src)
    assert len(hits) == 1, "anchor 3 (the process_queue call, NDP:1673) not found exactly once"
    src = DFS_ANCHOR_RE.sub(lambda m: DFS_NEW, src, count=1)
    return src


def main(argv):
    if len(argv) != 3:
        sys.stderr.write(__doc__)
        return 2
    with open(argv[1], encoding="utf-8", errors="surrogateescape") as f:
        src = f.read()
    with open(argv[2], "w", encoding="utf-8", errors="surrogateescape") as f:
        f.write(patch(src))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
