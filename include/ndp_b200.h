/* include/ndp_b200.h -- C ABI of the B200-native NDP hot path (BFS granulation -> DFS).
 *
 * The reference (GridSAT NDP 5.6.7) has no plugin/FFI layer; its hot path is entered at two
 * C++ call sites inside one translation unit.  Each entry point below names the reference
 * interface it replaces (file:line into the reference tree; NDP = NDP-5_6_7.cpp,
 * CSP = ClauseSetPool.hpp).  Plain pointers and sizes only; no exceptions cross this boundary;
 * every function returns an ndp_status and leaves a message for ndp_last_error().
 *
 * Clause arrays are n x 3 row-major int32, exactly the layout flattenClauseSet produces
 * (NDP:1207-1215): literal 0 means "already false / padding" (CSP:37-39); a unit clause x is
 * {0,0,x} (NDP:632-634).  Variable ids are the caller's: any non-zero int32 except INT32_MIN, as sparse
 * as the caller likes (Clause3 is `int l[3]`, CSP:37-39).  The kernels pack literals in 16-bit codes, so a
 * CNF whose ids do not fit is renumbered inside the library (only variables that occur need a code) and
 * every literal that leaves the library -- frontier clauses, choices, models -- is mapped back; the limit
 * is 32766 DISTINCT variables (a 64-bit factoring CNF has 5952).
 *
 * There is no CPU fallback: every compute entry point fails with NDP_E_CUDA when no sm_100
 * device (or no CUDA driver) is present.
 */
#ifndef NDP_B200_H
#define NDP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ndp_status {
    NDP_OK = 0,
    NDP_E_INVALID = 1,      /* bad argument (null pointer, index out of range, ...) */
    NDP_E_UNSUPPORTED = 2,  /* input outside what the kernels hold (> 32766 DISTINCT variables, > 2^24 clauses, ...) */
    NDP_E_CUDA = 3,         /* CUDA runtime error / no usable device */
    NDP_E_NOMEM = 4         /* device arena or host allocation failed */
} ndp_status;

/* verdict bytes written by the DFS entry points */
enum { NDP_UNSAT = 0, NDP_SAT = 1, NDP_NOT_RUN = 2 };

/* Opaque frontier: the queue<pair<ClauseSet, vector<int>>> that Satisfy_iterative_BFS returns
 * (NDP:881-882, 943).  Owned by the library; clause sets live in a device arena of packed
 * clause slots (the device-side replacement of ClauseSetPool, CSP:43-69). */
typedef struct ndp_frontier ndp_frontier;

typedef struct ndp_stats {
    uint64_t nodes;          /* DFS nodes = Satisfy_iterative loop iterations (NDP:801) */
    uint64_t clause_evals;   /* sum of |A| over nodes that ran the resolution pass (NDP:664) */
    uint64_t rebuilds;       /* sibling re-materialisations (engine detail, not in the reference) */
    double kernel_ms;        /* device time of the DFS kernel(s), the piece split included, CUDA events */
    uint64_t h2d_bytes;      /* host->device bytes moved inside the call */
    uint64_t d2h_bytes;      /* device->host bytes moved inside the call */
    uint64_t launches;       /* kernels launched by the call */
} ndp_stats;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* Select the CUDA device used by subsequent calls from this thread's process (default 0).
 * Replaces nothing in the reference (MPI rank placement, NDP:1542-1545, is out of scope). */
int ndp_set_device(int device);
int ndp_device_count(int* count);
/* DFS engine: both visit the reference's nodes in the reference's order and return identical
 * results and counters.  SCAN streams each node's clause set (the reference's own data flow,
 * NDP:654-697, from shared memory); INDEX walks a per-variable occurrence index of the root
 * clause array and keeps a node as two bit-planes over its clause ids.  AUTO = INDEX where the
 * frontier has a root (it came from ndp_bfs, or ndp_frontier_attach_root accepted one), SCAN for a
 * bare ndp_frontier_from_host.  Env NDP_ENGINE=scan|index overrides AUTO. */
enum { NDP_ENGINE_AUTO = 0, NDP_ENGINE_SCAN = 1, NDP_ENGINE_INDEX = 2 };
int ndp_set_engine(int engine);
/* DFS work decomposition (index engine).  The reference's Satisfy_iterative is sequential inside
 * one subproblem (NDP:784-878) and a frontier can be narrower than the device (-d 12 yields ONE
 * subproblem: NDP:926,935).  A shard with fewer subproblems than the GPU has warps is therefore
 * first cut into PIECES -- subtrees of each subproblem's search tree, listed in the order the
 * reference's LIFO stack reaches them -- and the pieces are searched in parallel.  Results are
 * unchanged: same verdict per subproblem, the same (first-in-DFS-order) model, the same node and
 * clause-evaluation counts.  pieces = -1: automatic (default), 0: never split, > 0: split every
 * shard with fewer subproblems into about that many pieces.  Env NDP_SPLIT=<n>|off overrides -1. */
int ndp_set_split(long long pieces);
const char* ndp_last_error(void);
const char* ndp_version(void);

/* ---- BFS granulation --------------------------------------------------------------------
 * Replaces the call at NDP:1651:
 *   Satisfy_iterative_BFS(clauses, depth, override_max_iterations, iterations, max_queues,
 *                         world_rank, initial_queue_size)            (definition NDP:881-944)
 * Same argument meaning: max_iterations is the -d node-expansion budget (NDP:935), max_queues
 * the -q queue-size threshold tested before each pop (NDP:894), override_max_iterations what
 * -q sets (NDP:1512).  Outputs: *iterations (NDP:926), *task_count (second member of the
 * returned pair, NDP:943), *initial_queue_size with the reference's in/out rule "assign only
 * if it was 0" (NDP:939-941).  The frontier holds the same subproblems in the same order. */
int ndp_bfs(const int32_t* clauses3, size_t n_clauses, int max_iterations, int override_max_iterations,
            int max_queues, ndp_frontier** out, int* iterations, int* task_count, size_t* initial_queue_size);

/* BFS clause evaluations (sum of |A| over expanded nodes) and device time of the last ndp_bfs
 * that produced this frontier; 0 for frontiers built by ndp_frontier_from_host. */
int ndp_frontier_bfs_stats(const ndp_frontier* f, uint64_t* clause_evals, double* kernel_ms, uint64_t* launches);
/* host<->device bytes the BFS (or ndp_frontier_from_host) moved while building this frontier */
int ndp_frontier_bfs_bytes(const ndp_frontier* f, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

size_t ndp_frontier_size(const ndp_frontier* f);

/* Element k of the queue in FIFO order: the materialised clause set (zeros included) and the
 * choices that led to it -- what serializeBFSResults walks (NDP:487-510) and what the
 * dispatcher ships to a worker (NDP:1337-1346).  Returned pointers are library-owned and stay
 * valid until the next ndp_frontier_get on the same frontier or ndp_frontier_free. */
int ndp_frontier_get(ndp_frontier* f, size_t k, const int32_t** clauses3, size_t* n_clauses,
                     const int32_t** choices, size_t* n_choices);

/* Sizes only (no device transfer). */
int ndp_frontier_dims(const ndp_frontier* f, size_t k, size_t* n_clauses, size_t* n_choices);

/* Build a frontier from host data: the -r resume path, replacing deserializeBFSResults
 * (NDP:512-531).  clause_offsets/choice_offsets have count+1 entries (in clauses / in ints).
 * DFS then runs on exactly these clause sets, as the reference's workers do (NDP:1422-1424). */
int ndp_frontier_from_host(const int32_t* clauses3, const size_t* clause_offsets, const int32_t* choices,
                           const size_t* choice_offsets, size_t count, ndp_frontier** out);

/* Tell a frontier built by ndp_frontier_from_host which clause array it was granulated from (the
 * DIMACS file the reference still requires on -r, NDP:1584-1626).  The library checks, on the device,
 * that every element really is that root filtered by its choices (the equivalence ResolutionStep
 * guarantees, NDP:700-750) and then searches the frontier with the occurrence-index engine -- the same
 * speed and the same piece split as a frontier that came straight from ndp_bfs.  If the check fails
 * (another file, edited JSON) NDP_E_INVALID is returned, the frontier is left as it was and is searched
 * by the clause-streaming engine on exactly the clause sets it holds, as the reference does. */
int ndp_frontier_attach_root(ndp_frontier* f, const int32_t* clauses3, size_t n_clauses);

void ndp_frontier_free(ndp_frontier* f);

/* ---- DFS over the frontier --------------------------------------------------------------
 * Replaces process_queue's dispatch loop + the per-task call at NDP:1424,
 *   Satisfy_iterative(*clause_set, true)                             (definition NDP:784-878)
 * for subproblems first, first+stride, first+2*stride, ... (< ndp_frontier_size): one rank's
 * static interleaved shard (rank = first, world = stride); first=0,stride=1 is the whole
 * frontier.  verdict[] has ndp_frontier_size entries indexed by GLOBAL subproblem index; only
 * this shard's entries are written (others keep the caller's value).
 *
 * early_exit = 0: every subproblem of the shard is resolved (exhaustive).
 * early_exit = 1: the reference's "first solution ends the job" (NDP:1297-1303, 1436-1447) made
 *   deterministic: the winner is the LOWEST-index SAT subproblem; subproblems above a known
 *   SAT index are abandoned (NDP_NOT_RUN).  *exit_flag (may be NULL) is an in/out hint shared
 *   between shards: on entry the lowest SAT index known so far (or SIZE_MAX), on return the
 *   lowest after this shard.
 * winner: global index of the lowest SAT subproblem of this shard, or SIZE_MAX.
 * model/model_len: choices_k ++ DFS choices of the winner -- final_choices_i of NDP:1430-1431;
 *   model must hold model_cap ints.  One case returns a winner WITHOUT a model (model_len = 0): a lower SAT
 *   index was already known when this shard's winner finished (the *exit_flag hint on entry, or a peer's exit
 *   word) -- the job's model then comes from the shard that owns that lower index, and this shard's own SAT
 *   subproblem is only reported in verdict[] / winner.
 * nodes/evals (may be NULL): per-subproblem counters, ndp_frontier_size entries each. */
int ndp_dfs_shard(ndp_frontier* f, size_t first, size_t stride, int early_exit, size_t* exit_flag,
                  uint8_t* verdict, size_t* winner, int32_t* model, size_t model_cap, size_t* model_len,
                  uint64_t* nodes, uint64_t* evals, ndp_stats* stats);

/* ---- early exit across processes (one process per GPU) -------------------------------------
 * The reference ends the job when the first worker reports a model (TAG_SOLUTION_FOUND, NDP:1297-1303,
 * 1436): rank 0 then tells every worker to stop.  With one process per GPU and no dispatcher, the same
 * signal is one 8-byte word per device holding the lowest SAT subproblem index known so far: the kernel
 * that finds a model lowers its own word and every peer's (system-scope atomicMin over NVLink), and
 * every kernel polls its own.  A word is created on the calling process's current device, exported as
 * a 64-byte CUDA IPC handle (ship it with any transport: bench.py uses one NCCL all-gather), and
 * imported by the other processes. */
typedef struct ndp_exit_word ndp_exit_word;
enum { NDP_EXIT_HANDLE_BYTES = 64 };
int ndp_exit_word_create(ndp_exit_word** word);
int ndp_exit_word_export(const ndp_exit_word* word, uint8_t handle[NDP_EXIT_HANDLE_BYTES]);
int ndp_exit_word_import(const uint8_t handle[NDP_EXIT_HANDLE_BYTES], ndp_exit_word** peer);
int ndp_exit_word_reset(ndp_exit_word* word);                       /* back to "no model known" (own words only) */
int ndp_exit_word_read(const ndp_exit_word* word, size_t* lowest);  /* SIZE_MAX: none */
void ndp_exit_word_free(ndp_exit_word* word);

/* ndp_dfs_shard with the early-exit words above instead of the in/out hint: `own` is this process's word,
 * peers[0..n_peers) the imported words of the other ranks (n_peers <= 7).  Subproblems above the lowest
 * SAT index seen in `own` are abandoned (NDP_NOT_RUN), those below always finish, so the winner of the
 * job -- the minimum of the ranks' winners -- is the same as on one GPU. */
int ndp_dfs_shard_shared(ndp_frontier* f, size_t first, size_t stride, int early_exit, ndp_exit_word* own,
                         ndp_exit_word* const* peers, int n_peers, uint8_t* verdict, size_t* winner,
                         int32_t* model, size_t model_cap, size_t* model_len, uint64_t* nodes, uint64_t* evals,
                         ndp_stats* stats);

/* ---- piece-level sharing between processes (one process per GPU) ------------------------------
 * The reference's dispatcher hands out one task per request, so no worker idles while tasks remain
 * (NDP:1322-1346).  A static k mod world split of SUBPROBLEMS cannot do that when the frontier is narrow
 * or its subproblems are heavy-tailed; these three calls share out PIECES instead.  Every rank holds the
 * whole frontier (each ran the deterministic BFS, as every reference rank does, NDP:1651) and makes the
 * SAME DFS-order cut of it into pieces (ndp_set_split policy, sized for world devices); rank r searches
 * pieces r, r + world, ... -- hundreds to tens of thousands of near-equal units per rank, so the ranks finish together.
 * Results are per-piece and fold back per subproblem with two tiny collectives the CALLER runs (the
 * library holds no communicator):
 *   ndp_dfs_pieces_begin  cut + search this rank's pieces.  first_model[k] (ndp_frontier_size entries) =
 *                         lowest piece OF THIS RANK holding a model of subproblem k, UINT32_MAX if none.
 *                         own / peers: early-exit words as in ndp_dfs_shard_shared (may be NULL / 0); inside
 *                         a piece run a word holds the lowest PIECE index with a model.
 *   -> MIN all-reduce of first_model over the ranks
 *   ndp_dfs_pieces_fold   this rank's PARTIAL results given the job-wide first_model: verdict[k] (combine
 *                         with MAX: UNSAT 0 < SAT 1 < NOT_RUN 2; SAT is reported by every rank), nodes[k] /
 *                         evals[k] (combine with SUM: pieces up to the first model, the reference's counts),
 *                         winner = lowest SAT subproblem (the same on every rank), and -- only on the rank that
 *                         searched the winning piece, model_len = 0 elsewhere -- the model choices_k ++ DFS
 *                         choices (final_choices_i of NDP:1430-1431).  stats->nodes / clause_evals: this
 *                         rank's sums.
 *   ndp_dfs_pieces_end    frees the run.
 * A frontier wide enough for world devices is not cut: the pieces are the subproblems and the calls
 * reduce to ndp_dfs_shard_shared(first = rank, stride = world).  world = 1 is valid. */
typedef struct ndp_piece_run ndp_piece_run;
int ndp_dfs_pieces_begin(ndp_frontier* f, size_t rank, size_t world, int early_exit, ndp_exit_word* own,
                         ndp_exit_word* const* peers, int n_peers, ndp_piece_run** run, uint32_t* first_model,
                         ndp_stats* stats);
int ndp_dfs_pieces_fold(ndp_piece_run* run, const uint32_t* first_model, uint8_t* verdict, uint64_t* nodes,
                        uint64_t* evals, size_t* winner, int32_t* model, size_t model_cap, size_t* model_len,
                        ndp_stats* stats);
void ndp_dfs_pieces_end(ndp_piece_run* run);

/* Progress of the DFS call(s) in flight in this process, for the reference's 1 Hz status line
 * "DFS time: .. seconds  -  Remaining Queue Size: ..  -  Active Workers: .." (NDP:1271-1288).  May be called
 * from any host thread while another one is inside ndp_dfs_*: units (subproblems, or the pieces they were
 * cut into) handed to warps so far / in total, summed over the devices that are searching right now. */
int ndp_dfs_progress(size_t* units_pulled, size_t* units_total, int* devices_active);

/* Whole frontier on n_gpus devices of this process (static interleaved split k -> k mod n_gpus,
 * one host thread per device, peer-visible early-exit flag).  n_gpus <= 0 means 1.  A frontier too
 * narrow for the devices (-d 12, -d 20: ONE subproblem) is first cut into DFS-order pieces on its home
 * device and the pieces are what the devices share out; results are folded back per subproblem. */
int ndp_dfs_batch(ndp_frontier* f, int n_gpus, int early_exit, uint8_t* verdict, size_t* winner,
                  int32_t* model, size_t model_cap, size_t* model_len, ndp_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* NDP_B200_H */
