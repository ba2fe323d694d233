/* oracle/ndp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference hot path (GridSAT NDP 5.6.7), used solely as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The
 * product (ndp-5_b200/) must never include, link or call anything declared here.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the pin is
 * the reference itself: oracle/_ref/libndp_ref.so is the unmodified reference TU compiled in
 * this container (oracle/Makefile, oracle/ref_harness.cpp); tests/test_oracle_vs_ref.py
 * compares every function below against it, and tests/golden/ holds outputs the reference
 * produced (tools/make_golden.py) so the pin travels to machines without /root/reference.
 *
 * Clause arrays are n x 3 row-major int32 exactly as flattenClauseSet lays them out
 * (NDP-5_6_7.cpp:1207-1215); literal 0 = "already false / padding" (ClauseSetPool.hpp:37-39).
 */
#ifndef NDP_ORACLE_H
#define NDP_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_frontier orc_frontier;

/* parseDimacsString, NDP-5_6_7.cpp:619-642.  Returns the clause count; fills out3 if it fits. */
size_t orc_parse_dimacs(const char* text, int32_t* out3, size_t cap_clauses);

/* choice, NDP-5_6_7.cpp:753-781 */
int orc_choice(const int32_t* c3, size_t n);

/* ResolutionStep (NDP:700-750) / ResolutionStepWithConflict (NDP:654-697).
 * out_la/out_ra hold 3*n ints each.  conflict_* are -1 when with_conflict == 0. */
void orc_resolution_step(const int32_t* c3, size_t n, int i, int with_conflict, int32_t* out_la, size_t* n_la,
                         int32_t* out_ra, size_t* n_ra, int* conflict_la, int* conflict_ra);

/* Satisfy_iterative_BFS, NDP-5_6_7.cpp:881-944.  clause_evals (may be NULL) receives the sum of
 * |A| over expanded nodes, the work unit of SURVEY.md §8(d). */
orc_frontier* orc_bfs(const int32_t* c3, size_t n, int max_iterations, int override_max_iterations, int max_queues,
                      int* iterations, int* task_count, size_t* initial_queue_size, uint64_t* clause_evals);
size_t orc_frontier_size(const orc_frontier* f);
size_t orc_frontier_nclauses(const orc_frontier* f, size_t k);
size_t orc_frontier_nchoices(const orc_frontier* f, size_t k);
void orc_frontier_copy(const orc_frontier* f, size_t k, int32_t* clauses3, int32_t* choices);
void orc_frontier_free(orc_frontier* f);

/* Satisfy_iterative, NDP-5_6_7.cpp:784-878.  Returns the number of models found (0 => UNSAT;
 * with first_assignment != 0 at most 1).  The first model (decision-ordered signed literals)
 * is copied to model[0..cap).  nodes = loop iterations that reached choice(); clause_evals =
 * sum of |A| over nodes that ran ResolutionStepWithConflict. */
int orc_dfs(const int32_t* c3, size_t n, int first_assignment, int32_t* model, size_t cap, size_t* model_len,
            uint64_t* nodes, uint64_t* clause_evals);

/* design statistics of the last orc_dfs call; see ndp_oracle.c */
const uint64_t* orc_last_dfs_stats(void);

#ifdef __cplusplus
}
#endif
#endif
