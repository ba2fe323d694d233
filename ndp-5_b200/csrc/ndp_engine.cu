// ndp_engine.cu -- host side of the C ABI (include/ndp_b200.h): device arena, BFS level
// driver with the reference's exact -d/-q stop rules, DFS launch, result gather.
// There is no CPU fallback in this file: every compute entry point needs a CUDA device.
#include "../../include/ndp_b200.h"
#include "ndp_kernels.cuh"
#include "ndp_index.cuh"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace ndp;

// ------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;
thread_local int g_device = 0;
int g_engine = NDP_ENGINE_AUTO;   // process-wide; ndp_set_engine / env NDP_ENGINE

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? NDP_E_NOMEM : NDP_E_CUDA,                      \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                             \
    } while (0)

struct Entry {        // one frontier element
    int buf;          // arena buffer holding its clause slot
    uint32_t slot;
    int32_t n;
    int32_t tree;     // index into the choice tree (BFS frontiers) or -1
};

// What choice() (NDP-5_6_7.cpp:753-781) says about a host clause array: same encoding as SetInfo.
void host_setinfo(const int32_t* c3, size_t n, int& kind, int& lit) {
    if (n == 0) { kind = K_EMPTY; lit = 0; return; }
    int lit2 = 0, lit1 = 0;
    for (size_t k = 0; k < n && !lit2; ++k) {
        int zc = 0, last = 0;
        for (int j = 0; j < 3; ++j) { if (c3[3 * k + j] == 0) ++zc; else last = c3[3 * k + j]; }
        if (zc == 2) lit2 = last;
        else if (zc == 1 && !lit1) lit1 = last;
    }
    if (lit2) { kind = K_UNIT; lit = lit2; return; }
    kind = K_DECIDE;
    int v = lit1 ? lit1 : c3[0];
    lit = v < 0 ? -v : v;
}

}  // namespace

struct ndp_frontier {
    int device = 0;
    pclause* buf[2] = {nullptr, nullptr};   // the arena: two buffers of fixed-stride clause slots
    size_t cap_slots[2] = {0, 0};
    size_t stride = 0;                      // clauses per slot
    int max_var = 0;
    std::vector<Entry> entries;
    // choices: either a parent-pointer tree (BFS) or flat arrays (resume)
    std::vector<int32_t> tree_parent, tree_lit;
    std::vector<int32_t> flat_choices;
    std::vector<size_t> flat_off;
    // BFS stats
    uint64_t bfs_evals = 0, bfs_launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    double bfs_kernel_ms = 0;
    // scratch for ndp_frontier_get
    std::vector<int32_t> get_clauses, get_choices;
    std::vector<pclause> get_packed;
    // per-device DFS descriptors (built lazily): index = device ordinal position in `replicas`
    struct Replica {
        int device = -1;
        size_t first = 0, stride = 1;        // the shard this replica serves
        pclause* shard_buf = nullptr;        // private copy of the shard's slots (other devices only)
        const pclause** d_base = nullptr;    // [n_local]
        int32_t* d_n = nullptr;              // [n_local]
        size_t n_local = 0;
        int n_max = 0;
    };
    std::vector<Replica> replicas;
    // index engine (BFS-born frontiers only): host copy of the root, device copy of the choice tree
    bool has_root = false;
    std::vector<int32_t> root_c3;
    int32_t *d_tree_parent = nullptr, *d_tree_lit = nullptr;   // on `device`
    struct IndexReplica {
        int device = -1;
        size_t first = 0, stride = 1, n_local = 0;
        pclause* d_root = nullptr;
        unsigned *d_occ_off = nullptr, *d_occ = nullptr;
        unsigned* d_prefix_tables = nullptr;
        int table_words = 0, n_zero3 = 0, first_zero3 = -1;
    };
    std::vector<IndexReplica> index_replicas;

    void choices_of(size_t k, std::vector<int32_t>& out) const {
        out.clear();
        const Entry& e = entries[k];
        if (e.tree >= 0) {
            for (int32_t t = e.tree; t > 0; t = tree_parent[t]) out.push_back(tree_lit[t]);
            std::reverse(out.begin(), out.end());
        } else if (!flat_off.empty()) {
            out.assign(flat_choices.begin() + flat_off[k], flat_choices.begin() + flat_off[k + 1]);
        }
    }
    size_t nchoices_of(size_t k) const {
        const Entry& e = entries[k];
        if (e.tree >= 0) {
            size_t c = 0;
            for (int32_t t = e.tree; t > 0; t = tree_parent[t]) ++c;
            return c;
        }
        return flat_off.empty() ? 0 : flat_off[k + 1] - flat_off[k];
    }
    void free_replicas() {
        for (auto& r : replicas) {
            cudaSetDevice(r.device);
            if (r.shard_buf) cudaFree(r.shard_buf);
            if (r.d_base) cudaFree((void*)r.d_base);
            if (r.d_n) cudaFree(r.d_n);
        }
        replicas.clear();
    }
    ~ndp_frontier() {
        free_replicas();
        for (auto& r : index_replicas) {
            cudaSetDevice(r.device);
            if (r.d_root) cudaFree(r.d_root);
            if (r.d_occ_off) cudaFree(r.d_occ_off);
            if (r.d_occ) cudaFree(r.d_occ);
            if (r.d_prefix_tables) cudaFree(r.d_prefix_tables);
        }
        cudaSetDevice(device);
        if (d_tree_parent) cudaFree(d_tree_parent);
        if (d_tree_lit) cudaFree(d_tree_lit);
        for (int b = 0; b < 2; ++b)
            if (buf[b]) cudaFree(buf[b]);
    }
};

namespace {

int check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(NDP_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                    "); the NDP hot path has no CPU fallback");
    if (g_device >= n) return fail(NDP_E_INVALID, "selected device does not exist");
    CUDA_TRY(cudaSetDevice(g_device));
    return NDP_OK;
}

int ensure_buf(ndp_frontier* f, int b, size_t slots) {
    if (f->cap_slots[b] >= slots) return NDP_OK;
    if (f->buf[b]) { CUDA_TRY(cudaFree(f->buf[b])); f->buf[b] = nullptr; f->cap_slots[b] = 0; }
    size_t want = std::max(slots, (size_t)64);
    CUDA_TRY(cudaMalloc(&f->buf[b], want * f->stride * sizeof(pclause)));
    f->cap_slots[b] = want;
    return NDP_OK;
}

// Device-resident bookkeeping of the BFS level loop (freed on scope exit).
struct BfsLevelState {
    BfsNode* nodes[2] = {nullptr, nullptr};
    int32_t* node_tree[2] = {nullptr, nullptr};
    BfsChild* children = nullptr;
    int32_t *tree_parent = nullptr, *tree_lit = nullptr;
    BfsLevelTotals* totals = nullptr;
    size_t node_cap = 0, tree_cap = 0, tree_count = 0;
    int cur = 0;

    template <class T> static int grow(T*& p, size_t old_n, size_t new_n) {
        T* q = nullptr;
        CUDA_TRY(cudaMalloc(&q, new_n * sizeof(T)));
        if (p && old_n) CUDA_TRY(cudaMemcpy(q, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice));
        if (p) cudaFree(p);
        p = q;
        return NDP_OK;
    }
    int reserve_nodes(size_t want) {
        if (!totals) CUDA_TRY(cudaMalloc(&totals, sizeof(BfsLevelTotals)));
        if (want <= node_cap) return NDP_OK;
        size_t cap = std::max(want, std::max(node_cap * 2, (size_t)4096));
        int rc;
        for (int b = 0; b < 2; ++b) {
            if ((rc = grow(nodes[b], node_cap, cap))) return rc;
            if ((rc = grow(node_tree[b], node_cap, cap))) return rc;
        }
        if (children) cudaFree(children);
        children = nullptr;
        CUDA_TRY(cudaMalloc(&children, cap * sizeof(BfsChild)));
        node_cap = cap;
        return NDP_OK;
    }
    int reserve_tree(size_t want) {
        if (want <= tree_cap) return NDP_OK;
        size_t cap = std::max(want, std::max(tree_cap * 2, (size_t)65536));
        int rc;
        if ((rc = grow(tree_parent, tree_count, cap))) return rc;
        if ((rc = grow(tree_lit, tree_count, cap))) return rc;
        tree_cap = cap;
        return NDP_OK;
    }
    int download_level(size_t m, std::vector<BfsNode>& h_nodes, std::vector<int32_t>& h_tree, uint64_t& d2h) {
        h_nodes.resize(m);
        h_tree.resize(m);
        CUDA_TRY(cudaMemcpy(h_nodes.data(), nodes[cur], m * sizeof(BfsNode), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_tree.data(), node_tree[cur], m * sizeof(int32_t), cudaMemcpyDeviceToHost));
        d2h += m * (sizeof(BfsNode) + sizeof(int32_t));
        return NDP_OK;
    }
    ~BfsLevelState() {
        for (int b = 0; b < 2; ++b) { if (nodes[b]) cudaFree(nodes[b]); if (node_tree[b]) cudaFree(node_tree[b]); }
        if (children) cudaFree(children);
        if (tree_parent) cudaFree(tree_parent);
        if (tree_lit) cudaFree(tree_lit);
        if (totals) cudaFree(totals);
    }
};

// pack an int32 n x 3 clause array; returns false if a literal does not fit 16 bits
bool pack_host(const int32_t* c3, size_t n, pclause* out, int& max_var) {
    for (size_t k = 0; k < n; ++k) {
        for (int j = 0; j < 3; ++j) {
            int l = c3[3 * k + j];
            if (l > kMaxVar || l < -kMaxVar) return false;
            int v = l < 0 ? -l : l;
            if (v > max_var) max_var = v;
        }
        out[k] = pack_clause(c3[3 * k], c3[3 * k + 1], c3[3 * k + 2]);
    }
    return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" {

const char* ndp_last_error(void) { return g_err.c_str(); }
const char* ndp_version(void) { return "ndp-b200 0.1 (NDP 5.6.7 hot path, sm_100a)"; }

int ndp_set_device(int device) {
    if (device < 0) return fail(NDP_E_INVALID, "negative device ordinal");
    g_device = device;
    return NDP_OK;
}
int ndp_set_engine(int engine) {
    if (engine != NDP_ENGINE_AUTO && engine != NDP_ENGINE_SCAN && engine != NDP_ENGINE_INDEX)
        return fail(NDP_E_INVALID, "unknown engine");
    g_engine = engine;
    return NDP_OK;
}
int ndp_device_count(int* count) {
    if (!count) return fail(NDP_E_INVALID, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(NDP_E_CUDA, cudaGetErrorString(e)); }
    *count = n;
    return NDP_OK;
}

int ndp_bfs(const int32_t* clauses3, size_t n_clauses, int max_iterations, int override_max_iterations,
            int max_queues, ndp_frontier** out, int* iterations_out, int* task_count_out,
            size_t* initial_queue_size) {
    if (!out) return fail(NDP_E_INVALID, "out is null");
    *out = nullptr;
    if (!clauses3 && n_clauses) return fail(NDP_E_INVALID, "clauses3 is null");
    if (n_clauses > (size_t)INT_MAX / 4) return fail(NDP_E_UNSUPPORTED, "too many clauses");
    int rc = check_device();
    if (rc) return rc;

    std::unique_ptr<ndp_frontier> f(new ndp_frontier);
    f->device = g_device;
    f->stride = std::max(n_clauses, (size_t)1);
    std::vector<pclause> packed(f->stride);
    if (!pack_host(clauses3, n_clauses, packed.data(), f->max_var))
        return fail(NDP_E_UNSUPPORTED, "literal magnitude exceeds 32766 (16-bit clause packing)");
    if ((rc = ensure_buf(f.get(), 0, 64))) return rc;
    CUDA_TRY(cudaMemcpy(f->buf[0], packed.data(), n_clauses * sizeof(pclause), cudaMemcpyHostToDevice));
    f->h2d_bytes += n_clauses * sizeof(pclause);

    const int D = max_iterations, Q = max_queues;
    const bool ovr = override_max_iterations != 0;
    long long it = 0, task_count = 1;
    int cur_buf = 0;
    size_t m = 1;

    // Level state lives on the device: node list (slot, n, kind, lit), each node's index in the
    // choice tree, the tree itself (parent, literal).  The host reads back one 32-byte totals
    // record per level, and the full per-child metadata only for levels in which a stop test
    // can fire.
    BfsLevelState st;
    int status = NDP_OK;
    {
        BfsNode root{0, (int32_t)n_clauses, 0, 0};
        host_setinfo(clauses3, n_clauses, root.kind, root.lit);
        if ((status = st.reserve_nodes(1)) || (status = st.reserve_tree(1))) return status;
        int32_t zero = 0, minus1 = -1;
        CUDA_TRY(cudaMemcpy(st.nodes[0], &root, sizeof root, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(st.node_tree[0], &zero, 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(st.tree_parent, &minus1, 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(st.tree_lit, &zero, 4, cudaMemcpyHostToDevice));
        st.tree_count = 1;
        f->h2d_bytes += sizeof root + 12;
    }
    std::vector<cudaEvent_t> evs;          // one (start, stop) pair per level, summed at the end

    std::vector<BfsNode> h_nodes;          // current level, only materialised on the host when needed
    std::vector<int32_t> h_node_tree;
    std::vector<BfsChild> h_children;
    std::vector<Entry> final_entries;
    struct Pushed { int32_t parent_tree, lit; };
    std::vector<Pushed> host_tree_tail;    // tree entries created by the host walk of the stop level
    bool done = false;

    // Level-synchronous restatement of the FIFO loop NDP:893-937: a FIFO
    // queue only ever holds the unexpanded rest of level L followed by the level-L+1 children
    // pushed so far, so walking one level in order with the running queue size `s` reproduces
    // every stop test at the exact pop it would fire.
    while (!done) {
        if (m == 0) break;                                           // queue empty (NDP:893)
        bool stop_at_top = (Q > 0 && m >= (size_t)Q) || (D == -1 && !ovr);   // NDP:894-897 before the first pop
        size_t need = m;
        if (!stop_at_top && Q <= 0) {                                // -d budget known in advance
            long long R = (long long)D - it;
            need = (size_t)std::min<long long>((long long)m, std::max<long long>(R, 1));
        }
        // can any stop test fire inside this level?  -q: the queue never exceeds m + pushes <= 2m;
        // -d: the level is consumed whole iff it + m < D.
        const bool safe = !stop_at_top && (Q > 0 ? 2 * m < (size_t)Q : (it + (long long)m < (long long)D));
        if (stop_at_top) {
            if ((status = st.download_level(m, h_nodes, h_node_tree, f->d2h_bytes))) break;
            for (size_t j = 0; j < m; ++j) final_entries.push_back(Entry{cur_buf, h_nodes[j].slot, h_nodes[j].n, h_node_tree[j]});
            break;
        }
        {
            // grow the (dead) target buffer geometrically, but never past what the stop rule
            // allows a level to reach: width <= Q under -q, <= D+1 under -d
            size_t want = 2 * need;
            if (f->cap_slots[cur_buf ^ 1] < want) {
                size_t bound = SIZE_MAX;
                if (Q > 0) bound = 2 * (size_t)Q;
                else if (D > 0) bound = 2 * ((size_t)D + 1);
                want = std::min(std::max(want, f->cap_slots[cur_buf ^ 1] * 2), std::max(bound, want));
            }
            status = ensure_buf(f.get(), cur_buf ^ 1, want);
            if (status == NDP_E_NOMEM && want > 2 * need) {   // head-room did not fit: take the exact size
                cudaGetLastError();
                status = ensure_buf(f.get(), cur_buf ^ 1, 2 * need);
            }
            if (status) break;
        }
        if ((status = st.reserve_nodes(2 * need)) || (status = st.reserve_tree(st.tree_count + 2 * need))) break;
        const int warps_per_block = 8;
        int grid = (int)std::min<size_t>((need + warps_per_block - 1) / warps_per_block, (size_t)148 * 16);
        cudaEvent_t ea, eb;
        cudaEventCreate(&ea);
        cudaEventCreate(&eb);
        evs.push_back(ea);
        evs.push_back(eb);
        cudaEventRecord(ea, 0);
        bfs_expand_kernel<<<grid, warps_per_block * 32>>>(f->buf[cur_buf], f->buf[cur_buf ^ 1], f->stride,
                                                          st.nodes[st.cur], (int)need, st.children);
        bfs_compact_kernel<<<1, 1024>>>(st.nodes[st.cur], st.node_tree[st.cur], st.children, (int)need,
                                        st.nodes[st.cur ^ 1], st.node_tree[st.cur ^ 1], st.tree_parent, st.tree_lit,
                                        (int)st.tree_count, st.totals);
        cudaEventRecord(eb, 0);
        f->bfs_launches += 2;
        BfsLevelTotals tot;
        cudaError_t ce = cudaMemcpy(&tot, st.totals, sizeof tot, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) { status = fail(NDP_E_CUDA, std::string("BFS level: ") + cudaGetErrorString(ce)); break; }
        f->d2h_bytes += sizeof tot;
        if (safe) {
            // every node of the level is expanded; only the totals matter
            it += (long long)tot.expanded;
            task_count += (long long)tot.m_next;
            f->bfs_evals += tot.evals;
            st.tree_count += tot.m_next;
            st.cur ^= 1;
            cur_buf ^= 1;
            m = tot.m_next;
            continue;
        }
        // a stop test can fire here: walk the level on the host with the exact rules
        if ((status = st.download_level(m, h_nodes, h_node_tree, f->d2h_bytes))) break;
        h_children.resize(2 * need);
        ce = cudaMemcpy(h_children.data(), st.children, 2 * need * sizeof(BfsChild), cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) { status = fail(NDP_E_CUDA, std::string("BFS level: ") + cudaGetErrorString(ce)); break; }
        f->d2h_bytes += 2 * need * sizeof(BfsChild);
        std::vector<Entry> next_entries;
        size_t s = m, E = 0;
        bool stopped = false;
        for (size_t j = 0; j < need; ++j) {
            if (Q > 0 && s >= (size_t)Q) { stopped = true; break; }  // NDP:894-895 before this pop
            const BfsNode& nd = h_nodes[j];
            const int var = nd.kind == K_UNIT ? std::abs(nd.lit) : nd.lit;
            E = j + 1;                                               // popped (NDP:898-899)
            s -= 1;
            if (nd.n == 0 || var == 0) continue;                     // choice()==0: `continue`, no iteration (NDP:901-902)
            f->bfs_evals += (uint64_t)nd.n;
            for (int b = 0; b < 2; ++b) {                            // LA then RA (NDP:904-925)
                const BfsChild& c = h_children[2 * j + b];
                if (c.n < 0) continue;
                int32_t tree = (int32_t)(st.tree_count + host_tree_tail.size());
                host_tree_tail.push_back(Pushed{h_node_tree[j], b == 0 ? var : -var});
                next_entries.push_back(Entry{cur_buf ^ 1, (uint32_t)(2 * j + b), c.n, tree});
                ++task_count;
                ++s;
            }
            ++it;                                                    // NDP:926
            if (Q <= 0 && it >= D) { stopped = true; break; }        // NDP:935-936
        }
        if (stopped || E < m) {
            // frontier = unexpanded rest of the current level, then the children pushed so far
            for (size_t j = E; j < m; ++j) final_entries.push_back(Entry{cur_buf, h_nodes[j].slot, h_nodes[j].n, h_node_tree[j]});
            final_entries.insert(final_entries.end(), next_entries.begin(), next_entries.end());
            done = true;
            break;
        }
        // no stop fired: the device-side compaction of this level is exactly what the walk produced
        host_tree_tail.clear();
        st.tree_count += tot.m_next;
        st.cur ^= 1;
        cur_buf ^= 1;
        m = tot.m_next;
    }
    cudaDeviceSynchronize();
    for (size_t e = 0; e + 1 < evs.size(); e += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, evs[e], evs[e + 1]) == cudaSuccess) f->bfs_kernel_ms += ms;
    }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    if (status) return status;

    // bring the choice tree home (parent pointers + literals), then the stop level's tail
    f->tree_parent.resize(st.tree_count);
    f->tree_lit.resize(st.tree_count);
    CUDA_TRY(cudaMemcpy(f->tree_parent.data(), st.tree_parent, st.tree_count * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(f->tree_lit.data(), st.tree_lit, st.tree_count * 4, cudaMemcpyDeviceToHost));
    f->d2h_bytes += st.tree_count * 8;
    for (const Pushed& p : host_tree_tail) { f->tree_parent.push_back(p.parent_tree); f->tree_lit.push_back(p.lit); }
    f->entries.swap(final_entries);
    // the index engine derives every subproblem from the root + its choices: keep the root and
    // leave the tree on the device (completed with the stop level's tail)
    f->root_c3.assign(clauses3, clauses3 + 3 * n_clauses);
    f->has_root = true;
    if (!host_tree_tail.empty()) {
        if ((status = st.reserve_tree(f->tree_parent.size()))) return status;
        CUDA_TRY(cudaMemcpy(st.tree_parent + st.tree_count, f->tree_parent.data() + st.tree_count,
                            host_tree_tail.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(st.tree_lit + st.tree_count, f->tree_lit.data() + st.tree_count,
                            host_tree_tail.size() * 4, cudaMemcpyHostToDevice));
        f->h2d_bytes += host_tree_tail.size() * 8;
    }
    f->d_tree_parent = st.tree_parent;
    f->d_tree_lit = st.tree_lit;
    st.tree_parent = nullptr;   // ownership moves to the frontier
    st.tree_lit = nullptr;

    if (iterations_out) *iterations_out = (int)it;
    if (task_count_out) *task_count_out = (int)task_count;
    if (initial_queue_size && *initial_queue_size == 0) *initial_queue_size = f->entries.size();  // NDP:939-941
    *out = f.release();
    return NDP_OK;
}

int ndp_frontier_bfs_stats(const ndp_frontier* f, uint64_t* clause_evals, double* kernel_ms, uint64_t* launches) {
    if (!f) return fail(NDP_E_INVALID, "frontier is null");
    if (clause_evals) *clause_evals = f->bfs_evals;
    if (kernel_ms) *kernel_ms = f->bfs_kernel_ms;
    if (launches) *launches = f->bfs_launches;
    return NDP_OK;
}

int ndp_frontier_bfs_bytes(const ndp_frontier* f, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!f) return fail(NDP_E_INVALID, "frontier is null");
    if (h2d_bytes) *h2d_bytes = f->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = f->d2h_bytes;
    return NDP_OK;
}

size_t ndp_frontier_size(const ndp_frontier* f) { return f ? f->entries.size() : 0; }

int ndp_frontier_dims(const ndp_frontier* f, size_t k, size_t* n_clauses, size_t* n_choices) {
    if (!f || k >= f->entries.size()) return fail(NDP_E_INVALID, "frontier index out of range");
    if (n_clauses) *n_clauses = (size_t)f->entries[k].n;
    if (n_choices) *n_choices = f->nchoices_of(k);
    return NDP_OK;
}

int ndp_frontier_get(ndp_frontier* f, size_t k, const int32_t** clauses3, size_t* n_clauses,
                     const int32_t** choices, size_t* n_choices) {
    if (!f || k >= f->entries.size()) return fail(NDP_E_INVALID, "frontier index out of range");
    const Entry& e = f->entries[k];
    if (clauses3 || n_clauses) {
        CUDA_TRY(cudaSetDevice(f->device));
        f->get_packed.resize((size_t)std::max(e.n, 1));
        CUDA_TRY(cudaMemcpy(f->get_packed.data(), f->buf[e.buf] + (size_t)e.slot * f->stride,
                            (size_t)e.n * sizeof(pclause), cudaMemcpyDeviceToHost));
        f->get_clauses.resize((size_t)e.n * 3 + 1);
        for (int32_t c = 0; c < e.n; ++c) {
            const pclause p = f->get_packed[c];
            int l0, l1, l2;
            unpack_clause(p, l0, l1, l2);
            f->get_clauses[3 * c] = l0;
            f->get_clauses[3 * c + 1] = l1;
            f->get_clauses[3 * c + 2] = l2;
        }
        if (clauses3) *clauses3 = f->get_clauses.data();
        if (n_clauses) *n_clauses = (size_t)e.n;
    }
    if (choices || n_choices) {
        f->choices_of(k, f->get_choices);
        if (f->get_choices.empty()) f->get_choices.reserve(1);
        if (choices) *choices = f->get_choices.data();
        if (n_choices) *n_choices = f->get_choices.size();
    }
    return NDP_OK;
}

int ndp_frontier_from_host(const int32_t* clauses3, const size_t* clause_offsets, const int32_t* choices,
                           const size_t* choice_offsets, size_t count, ndp_frontier** out) {
    if (!out) return fail(NDP_E_INVALID, "out is null");
    *out = nullptr;
    if (count && (!clause_offsets || !choice_offsets)) return fail(NDP_E_INVALID, "offset arrays are null");
    int rc = check_device();
    if (rc) return rc;
    std::unique_ptr<ndp_frontier> f(new ndp_frontier);
    f->device = g_device;
    size_t stride = 1;
    for (size_t k = 0; k < count; ++k) {
        if (clause_offsets[k + 1] < clause_offsets[k] || choice_offsets[k + 1] < choice_offsets[k])
            return fail(NDP_E_INVALID, "offset arrays must be non-decreasing");
        stride = std::max(stride, clause_offsets[k + 1] - clause_offsets[k]);
    }
    if (stride > (size_t)INT_MAX / 4) return fail(NDP_E_UNSUPPORTED, "clause set too large");
    f->stride = stride;
    if ((rc = ensure_buf(f.get(), 0, std::max(count, (size_t)1)))) return rc;
    // stage in bounded chunks so a multi-GB resume does not double its footprint on the host
    const size_t chunk_slots = std::max<size_t>(1, ((size_t)64 << 20) / (stride * sizeof(pclause)));
    std::vector<pclause> stage(chunk_slots * stride);
    for (size_t k0 = 0; k0 < count; k0 += chunk_slots) {
        size_t k1 = std::min(count, k0 + chunk_slots);
        for (size_t k = k0; k < k1; ++k) {
            size_t n = clause_offsets[k + 1] - clause_offsets[k];
            if (!pack_host(clauses3 + 3 * clause_offsets[k], n, stage.data() + (k - k0) * stride, f->max_var))
                return fail(NDP_E_UNSUPPORTED, "literal magnitude exceeds 32766 (16-bit clause packing)");
            f->entries.push_back(Entry{0, (uint32_t)k, (int32_t)n, -1});
        }
        CUDA_TRY(cudaMemcpy(f->buf[0] + k0 * stride, stage.data(), (k1 - k0) * stride * sizeof(pclause),
                            cudaMemcpyHostToDevice));
        f->h2d_bytes += (k1 - k0) * stride * sizeof(pclause);
    }
    f->flat_off.assign(choice_offsets, choice_offsets + count + 1);
    size_t base = count ? choice_offsets[0] : 0;
    for (auto& o : f->flat_off) o -= base;
    if (count) f->flat_choices.assign(choices + choice_offsets[0], choices + choice_offsets[count]);
    if (count == 0) f->flat_off.assign(1, 0);
    for (int32_t l : f->flat_choices) f->max_var = std::max(f->max_var, std::abs(l));
    *out = f.release();
    return NDP_OK;
}

void ndp_frontier_free(ndp_frontier* f) { delete f; }

}  // extern "C"

// ------------------------------------------------------------------------------------------
namespace {

struct DfsLaunchPlan {
    bool smem_w;
    int warps_per_cta, ctas_per_sm, grid, smem_bytes;
    DfsParams p;
};

// shared-memory layout of one warp and the launch shape that maximises resident warps
int plan_dfs(int device, int n_max, int max_var, DfsLaunchPlan& plan) {
    int sms = 0, smem_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    DfsParams& p = plan.p;
    p.w_cap = pad_up(std::max(n_max, 1));
    p.trail_cap = std::max(max_var, 1);
    p.table_words = (max_var + 1 + 15) / 16;
    p.pend_words = (p.trail_cap + 31) / 32;
    const size_t aux = align16((size_t)p.trail_cap * 2) + align16((size_t)p.table_words * 4) + align16((size_t)p.pend_words * 4);
    const size_t w_bytes = align16((size_t)p.w_cap * sizeof(pclause));
    plan.smem_w = w_bytes + aux <= (size_t)smem_optin;
    const size_t per_warp = (plan.smem_w ? w_bytes : 0) + aux;
    if (per_warp > (size_t)smem_optin)
        return fail(NDP_E_UNSUPPORTED, "variable count too large for the shared-memory trail");
    size_t off = plan.smem_w ? w_bytes : 0;
    p.off_trail = (int)off; off += align16((size_t)p.trail_cap * 2);
    p.off_table = (int)off; off += align16((size_t)p.table_words * 4);
    p.off_pend = (int)off;  off += align16((size_t)p.pend_words * 4);
    p.warp_bytes = (int)per_warp;
    // pick warps/CTA in {8,6,4,3,2,1} maximising resident warps per SM (<= 32 to keep registers free)
    int best_total = 0, best_w = 1, best_ctas = 1;
    const int sm_smem = 233472 - 1024;  // 228 KB per SM, 1 KB reserved per resident CTA
    for (int w : {8, 6, 4, 3, 2, 1}) {
        size_t cta = per_warp * w;
        if (cta > (size_t)smem_optin) continue;
        int ctas = (int)std::min<size_t>(32, (size_t)sm_smem / (cta + 1024));
        ctas = std::min(ctas, 32 / w);
        if (ctas < 1) continue;
        if (w * ctas > best_total) { best_total = w * ctas; best_w = w; best_ctas = ctas; }
    }
    plan.warps_per_cta = best_w;
    plan.ctas_per_sm = best_ctas;
    plan.smem_bytes = (int)(per_warp * best_w);
    plan.grid = sms * best_ctas;
    return NDP_OK;
}

int build_replica(ndp_frontier* f, int device, size_t first, size_t stride, ndp_frontier::Replica** out) {
    for (auto& r : f->replicas)
        if (r.device == device && r.first == first && r.stride == stride) { *out = &r; return NDP_OK; }
    ndp_frontier::Replica r;
    r.device = device;
    r.first = first;
    r.stride = stride;
    const size_t total = f->entries.size();
    r.n_local = first < total ? (total - first + stride - 1) / stride : 0;
    CUDA_TRY(cudaSetDevice(device));
    std::vector<const pclause*> h_base(std::max(r.n_local, (size_t)1));
    std::vector<int32_t> h_n(std::max(r.n_local, (size_t)1));
    if (device != f->device && r.n_local) {
        // private copy of this shard's slots on the other GPU, one peer copy per slot
        CUDA_TRY(cudaMalloc(&r.shard_buf, r.n_local * f->stride * sizeof(pclause)));
        for (size_t q = 0; q < r.n_local; ++q) {
            const Entry& e = f->entries[first + q * stride];
            CUDA_TRY(cudaMemcpyPeerAsync(r.shard_buf + q * f->stride, device,
                                         f->buf[e.buf] + (size_t)e.slot * f->stride, f->device,
                                         (size_t)e.n * sizeof(pclause), 0));
        }
        CUDA_TRY(cudaStreamSynchronize(0));
    }
    for (size_t q = 0; q < r.n_local; ++q) {
        const Entry& e = f->entries[first + q * stride];
        h_base[q] = r.shard_buf ? r.shard_buf + q * f->stride : f->buf[e.buf] + (size_t)e.slot * f->stride;
        h_n[q] = e.n;
        r.n_max = std::max(r.n_max, e.n);
    }
    CUDA_TRY(cudaMalloc((void**)&r.d_base, h_base.size() * sizeof(pclause*)));
    CUDA_TRY(cudaMalloc(&r.d_n, h_n.size() * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy((void*)r.d_base, h_base.data(), h_base.size() * sizeof(pclause*), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_n, h_n.data(), h_n.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    f->replicas.push_back(r);
    *out = &f->replicas.back();
    return NDP_OK;
}

struct DevBuf {  // tiny RAII for the per-call device outputs
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T* as() { return static_cast<T*>(p); }
};

// ---- index engine ---------------------------------------------------------------------------

bool engine_uses_index(const ndp_frontier* f) {
    int e = g_engine;
    if (e == NDP_ENGINE_AUTO) {
        const char* s = getenv("NDP_ENGINE");
        if (s && !strcmp(s, "scan")) e = NDP_ENGINE_SCAN;
        else if (s && !strcmp(s, "index")) e = NDP_ENGINE_INDEX;
    }
    if (!f->has_root) return false;   // explicit clause sets (resume): no common root to index
    return e != NDP_ENGINE_SCAN;
}

struct IndexLaunchPlan {
    int warps_per_cta, ctas_per_sm, grid, smem_bytes;
    IndexParams p;
};

int plan_index(int device, int n_root, int max_var, IndexLaunchPlan& plan) {
    int sms = 0, smem_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    IndexParams& p = plan.p;
    p.bm_words = (n_root + 31) / 32;
    p.table_words = (max_var + 1 + 15) / 16;
    p.trail_cap = std::max(max_var, 1);
    p.pend_words = (p.trail_cap + 31) / 32;
    size_t off = align16((size_t)3 * p.bm_words * 4);
    p.off_table = (int)off; off += align16((size_t)p.table_words * 4);
    p.off_pend = (int)off;  off += align16((size_t)p.pend_words * 4);
    p.warp_bytes = (int)std::max(off, (size_t)16);
    if ((size_t)p.warp_bytes > (size_t)smem_optin)
        return fail(NDP_E_UNSUPPORTED, "clause count too large for the shared-memory bitmaps");
    const int sm_smem = 233472;
    int best_total = 0, best_w = 1, best_ctas = 1;
    int cap_warps = 32;   // resident warps per SM; tunable for experiments
    if (const char* s = getenv("NDP_INDEX_WARPS_PER_SM")) cap_warps = std::max(1, std::min(64, atoi(s)));
    for (int w : {4, 2, 1}) {
        size_t cta = (size_t)p.warp_bytes * w;
        if (cta > (size_t)smem_optin) continue;
        int ctas = (int)std::min<size_t>(cap_warps / w, (size_t)sm_smem / (cta + 1024));
        if (ctas < 1) continue;
        if (w * ctas > best_total) { best_total = w * ctas; best_w = w; best_ctas = ctas; }
    }
    plan.warps_per_cta = best_w;
    plan.ctas_per_sm = best_ctas;
    plan.smem_bytes = p.warp_bytes * best_w;
    plan.grid = sms * best_ctas;
    return NDP_OK;
}

int build_index_replica(ndp_frontier* f, int device, size_t first, size_t stride, ndp_frontier::IndexReplica** out) {
    for (auto& r : f->index_replicas)
        if (r.device == device && r.first == first && r.stride == stride) { *out = &r; return NDP_OK; }
    ndp_frontier::IndexReplica r;
    r.device = device;
    r.first = first;
    r.stride = stride;
    const size_t total = f->entries.size();
    r.n_local = first < total ? (total - first + stride - 1) / stride : 0;
    const size_t n_root = f->root_c3.size() / 3;
    const int V = f->max_var;
    // root records + merged occurrence lists (c | n_pos << 24 | n_neg << 26), built on the host once
    std::vector<pclause> root(std::max(n_root, (size_t)1));
    std::vector<unsigned> off((size_t)V + 2, 0), occ;
    if (n_root >= (1u << 24)) return fail(NDP_E_UNSUPPORTED, "more than 2^24 root clauses");
    {
        std::vector<std::vector<unsigned>> per_var((size_t)V + 1);
        for (size_t c = 0; c < n_root; ++c) {
            const int32_t* l = &f->root_c3[3 * c];
            root[c] = pack_clause(l[0], l[1], l[2]);
            if (!l[0] && !l[1] && !l[2]) { ++r.n_zero3; if (r.first_zero3 < 0) r.first_zero3 = (int)c; }
            for (int j = 0; j < 3; ++j) {
                if (!l[j]) continue;
                bool seen = false;
                for (int i = 0; i < j; ++i) seen |= l[i] && std::abs(l[i]) == std::abs(l[j]);
                if (seen) continue;
                unsigned n_pos = 0, n_neg = 0;
                for (int i = 0; i < 3; ++i) {
                    if (l[i] == std::abs(l[j])) ++n_pos;
                    if (l[i] == -std::abs(l[j])) ++n_neg;
                }
                per_var[(size_t)std::abs(l[j])].push_back((unsigned)c | (n_pos << 24) | (n_neg << 26));
            }
        }
        for (int v = 0; v <= V; ++v) {
            off[v] = (unsigned)occ.size();
            occ.insert(occ.end(), per_var[v].begin(), per_var[v].end());
        }
        off[(size_t)V + 1] = (unsigned)occ.size();
    }
    r.table_words = (V + 1 + 15) / 16;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMalloc(&r.d_root, root.size() * sizeof(pclause)));
    CUDA_TRY(cudaMalloc(&r.d_occ_off, off.size() * 4));
    CUDA_TRY(cudaMalloc(&r.d_occ, std::max(occ.size(), (size_t)1) * 4));
    CUDA_TRY(cudaMemcpy(r.d_root, root.data(), root.size() * sizeof(pclause), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_occ_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    if (!occ.empty()) CUDA_TRY(cudaMemcpy(r.d_occ, occ.data(), occ.size() * 4, cudaMemcpyHostToDevice));
    f->h2d_bytes += root.size() * sizeof(pclause) + off.size() * 4 + occ.size() * 4;
    // each subproblem's choices as a 2-bit assignment table, derived on the device from the tree
    const size_t tbytes = std::max(r.n_local, (size_t)1) * r.table_words * 4;
    CUDA_TRY(cudaMalloc(&r.d_prefix_tables, tbytes));
    CUDA_TRY(cudaMemset(r.d_prefix_tables, 0, tbytes));
    if (r.n_local) {
        std::vector<int32_t> h_entry_tree(total);
        for (size_t k = 0; k < total; ++k) h_entry_tree[k] = f->entries[k].tree;
        DevBuf d_entry, d_tp, d_tl;
        CUDA_TRY(cudaMalloc(&d_entry.p, total * 4));
        CUDA_TRY(cudaMemcpy(d_entry.p, h_entry_tree.data(), total * 4, cudaMemcpyHostToDevice));
        f->h2d_bytes += total * 4;
        const int32_t *tp = f->d_tree_parent, *tl = f->d_tree_lit;
        if (device != f->device || !tp) {   // another GPU: ship the tree from the host copy
            const size_t tn = f->tree_parent.size();
            CUDA_TRY(cudaMalloc(&d_tp.p, tn * 4));
            CUDA_TRY(cudaMalloc(&d_tl.p, tn * 4));
            CUDA_TRY(cudaMemcpy(d_tp.p, f->tree_parent.data(), tn * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(d_tl.p, f->tree_lit.data(), tn * 4, cudaMemcpyHostToDevice));
            f->h2d_bytes += tn * 8;
            tp = d_tp.as<int32_t>();
            tl = d_tl.as<int32_t>();
        }
        const int threads = 128;
        const int blocks = (int)std::min<size_t>((r.n_local + threads - 1) / threads, 4096);
        prefix_table_kernel<<<blocks, threads>>>(d_entry.as<int32_t>(), tp, tl, first, stride, r.n_local,
                                                 r.table_words, r.d_prefix_tables);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
    }
    f->index_replicas.push_back(r);
    *out = &f->index_replicas.back();
    return NDP_OK;
}

// One shard on one device.  `best_dev`/`peers`: early-exit words (this device's + peers').
int dfs_on_device(ndp_frontier* f, int device, size_t first, size_t stride, int early_exit,
                  unsigned long long* shared_best, unsigned long long* const* peer_best, int n_peers,
                  unsigned long long best_init, uint8_t* verdict, size_t* winner, std::vector<int32_t>* model,
                  uint64_t* nodes, uint64_t* evals, ndp_stats* stats) {
    const bool use_index = engine_uses_index(f);
    int rc;
    ndp_frontier::Replica* rep = nullptr;
    ndp_frontier::IndexReplica* irep = nullptr;
    if (use_index) rc = build_index_replica(f, device, first, stride, &irep);
    else rc = build_replica(f, device, first, stride, &rep);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    const size_t nl = use_index ? irep->n_local : rep->n_local;
    if (winner) *winner = SIZE_MAX;
    if (nl == 0) return NDP_OK;
    if (nl > 0xfffffff0ull) return fail(NDP_E_UNSUPPORTED, "shard larger than the 32-bit work counter");

    DfsLaunchPlan plan;
    IndexLaunchPlan iplan;
    int n_warps, model_stride;
    if (use_index) {
        if ((rc = plan_index(device, (int)(f->root_c3.size() / 3), f->max_var, iplan))) return rc;
        n_warps = iplan.grid * iplan.warps_per_cta;
        model_stride = iplan.p.trail_cap;
    } else {
        if ((rc = plan_dfs(device, rep->n_max, f->max_var, plan))) return rc;
        n_warps = plan.grid * plan.warps_per_cta;
        plan.p.model_stride = plan.p.trail_cap;
        model_stride = plan.p.model_stride;
    }

    DevBuf b_verdict, b_nodes, b_evals, b_reb, b_next, b_best, b_msub, b_mlen, b_models, b_scratch;
    CUDA_TRY(cudaMalloc(&b_verdict.p, nl));
    CUDA_TRY(cudaMalloc(&b_nodes.p, nl * 8));
    CUDA_TRY(cudaMalloc(&b_evals.p, nl * 8));
    CUDA_TRY(cudaMalloc(&b_reb.p, nl * 4));
    CUDA_TRY(cudaMalloc(&b_next.p, 4));
    CUDA_TRY(cudaMalloc(&b_msub.p, (size_t)n_warps * 8));
    CUDA_TRY(cudaMalloc(&b_mlen.p, (size_t)n_warps * 4));
    CUDA_TRY(cudaMalloc(&b_models.p, (size_t)n_warps * model_stride * 4));
    if (use_index) CUDA_TRY(cudaMalloc(&b_scratch.p, (size_t)n_warps * model_stride * sizeof(short)));   // trails
    else if (!plan.smem_w) CUDA_TRY(cudaMalloc(&b_scratch.p, (size_t)n_warps * plan.p.w_cap * sizeof(pclause)));
    unsigned long long* d_best = shared_best;
    if (!d_best) {
        CUDA_TRY(cudaMalloc(&b_best.p, 8));
        d_best = b_best.as<unsigned long long>();
        CUDA_TRY(cudaMemcpy(d_best, &best_init, 8, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMemset(b_verdict.p, NDP_NOT_RUN, nl));
    CUDA_TRY(cudaMemset(b_next.p, 0, 4));
    CUDA_TRY(cudaMemset(b_msub.p, 0xff, (size_t)n_warps * 8));

    cudaEvent_t ev0, ev1;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    if (use_index) {
        IndexParams& p = iplan.p;
        p.root = irep->d_root;
        p.n_root = (int)(f->root_c3.size() / 3);
        p.occ_off = irep->d_occ_off;
        p.occ = irep->d_occ;
        p.prefix_tables = irep->d_prefix_tables;
        p.n_zero3 = irep->n_zero3;
        p.first_zero3 = irep->first_zero3;
        p.first = first; p.stride = stride; p.n_local = nl;
        p.next = b_next.as<unsigned int>();
        p.best = d_best;
        p.n_peers = 0;
        for (int d = 0; d < n_peers && d < 8; ++d)
            if (peer_best[d] != d_best) p.peer_best[p.n_peers++] = peer_best[d];
        p.early_exit = early_exit;
        p.verdict = b_verdict.as<uint8_t>();
        p.nodes = b_nodes.as<unsigned long long>();
        p.evals = b_evals.as<unsigned long long>();
        p.rebuilds = b_reb.as<unsigned int>();
        p.model_sub = b_msub.as<unsigned long long>();
        p.model_len = b_mlen.as<int32_t>();
        p.trails = b_scratch.as<short>();
        p.models = b_models.as<int32_t>();
        CUDA_TRY(cudaFuncSetAttribute(dfs_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, iplan.smem_bytes));
        {
            // leave L1 everything the resident CTAs do not need: the root index is read through it
            const size_t need = (size_t)iplan.ctas_per_sm * (iplan.smem_bytes + 1024);
            int pct = (int)std::min<size_t>(100, (need * 100 + 233471) / 233472);
            cudaFuncSetAttribute(dfs_index_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        CUDA_TRY(cudaEventRecord(ev0, 0));
        dfs_index_kernel<<<iplan.grid, iplan.warps_per_cta * 32, iplan.smem_bytes>>>(p);
    } else {
        DfsParams& p = plan.p;
        p.sub_base = rep->d_base;
        p.sub_n = rep->d_n;
        p.first = first; p.stride = stride; p.n_local = nl;
        p.next = b_next.as<unsigned int>();
        p.best = d_best;
        p.n_peers = 0;
        for (int d = 0; d < n_peers && d < 8; ++d)
            if (peer_best[d] != d_best) p.peer_best[p.n_peers++] = peer_best[d];
        p.early_exit = early_exit;
        p.verdict = b_verdict.as<uint8_t>();
        p.nodes = b_nodes.as<unsigned long long>();
        p.evals = b_evals.as<unsigned long long>();
        p.rebuilds = b_reb.as<unsigned int>();
        p.model_sub = b_msub.as<unsigned long long>();
        p.model_len = b_mlen.as<int32_t>();
        p.models = b_models.as<int32_t>();
        p.gscratch = b_scratch.as<pclause>();
        auto kern = plan.smem_w ? dfs_kernel<true> : dfs_kernel<false>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes));
        CUDA_TRY(cudaEventRecord(ev0, 0));
        kern<<<plan.grid, plan.warps_per_cta * 32, plan.smem_bytes>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev1, 0));
    CUDA_TRY(cudaEventSynchronize(ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);

    std::vector<uint8_t> h_v(nl);
    std::vector<unsigned long long> h_nodes(nl), h_evals(nl);
    std::vector<unsigned> h_reb(nl);
    CUDA_TRY(cudaMemcpy(h_v.data(), b_verdict.p, nl, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(h_nodes.data(), b_nodes.p, nl * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(h_evals.data(), b_evals.p, nl * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(h_reb.data(), b_reb.p, nl * 4, cudaMemcpyDeviceToHost));
    size_t win = SIZE_MAX;
    uint64_t tn = 0, te = 0, tr = 0;
    for (size_t q = 0; q < nl; ++q) {
        const size_t g = first + q * stride;
        if (verdict) verdict[g] = h_v[q];
        if (nodes) nodes[g] = h_nodes[q];
        if (evals) evals[g] = h_evals[q];
        tn += h_nodes[q]; te += h_evals[q]; tr += h_reb[q];
        if (h_v[q] == NDP_SAT && win == SIZE_MAX) win = g;
    }
    if (winner) *winner = win;
    if (win != SIZE_MAX && model) {
        std::vector<unsigned long long> h_msub(n_warps);
        std::vector<int32_t> h_mlen(n_warps);
        CUDA_TRY(cudaMemcpy(h_msub.data(), b_msub.p, (size_t)n_warps * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_mlen.data(), b_mlen.p, (size_t)n_warps * 4, cudaMemcpyDeviceToHost));
        model->clear();
        for (int w = 0; w < n_warps; ++w)
            if (h_msub[w] == (unsigned long long)win) {
                f->choices_of(win, *model);                       // choices_k ++ dfs choices (NDP:1430-1431)
                std::vector<int32_t> tail((size_t)h_mlen[w]);
                CUDA_TRY(cudaMemcpy(tail.data(), b_models.as<int32_t>() + (size_t)w * model_stride,
                                    tail.size() * 4, cudaMemcpyDeviceToHost));
                model->insert(model->end(), tail.begin(), tail.end());
                break;
            }
    }
    if (stats) {
        stats->h2d_bytes += 8 + 4;                                   // early-exit word + work counter
        stats->d2h_bytes += nl * (1 + 8 + 8 + 4) + (win != SIZE_MAX && model ? (size_t)n_warps * 12 + model->size() * 4 : 0);
        stats->nodes += tn;
        stats->clause_evals += te;
        stats->rebuilds += tr;
        stats->kernel_ms = std::max(stats->kernel_ms, (double)ms);
        stats->launches += 1;
    }
    return NDP_OK;
}

}  // namespace

extern "C" {

int ndp_dfs_shard(ndp_frontier* f, size_t first, size_t stride, int early_exit, size_t* exit_flag,
                  uint8_t* verdict, size_t* winner, int32_t* model, size_t model_cap, size_t* model_len,
                  uint64_t* nodes, uint64_t* evals, ndp_stats* stats) {
    if (!f) return fail(NDP_E_INVALID, "frontier is null");
    if (stride == 0) return fail(NDP_E_INVALID, "stride must be >= 1");
    int rc = check_device();
    if (rc) return rc;
    if (stats) memset(stats, 0, sizeof *stats);
    if (model_len) *model_len = 0;
    unsigned long long best_init = ~0ull;
    if (exit_flag && *exit_flag != SIZE_MAX) best_init = *exit_flag;
    std::vector<int32_t> m;
    size_t win = SIZE_MAX;
    rc = dfs_on_device(f, f->device, first, stride, early_exit, nullptr, nullptr, 0, best_init, verdict, &win,
                       &m, nodes, evals, stats);
    if (rc) return rc;
    if (winner) *winner = win;
    if (win != SIZE_MAX) {
        if (exit_flag && (*exit_flag == SIZE_MAX || win < *exit_flag)) *exit_flag = win;
        if (model_len) *model_len = m.size();
        if (model) {
            if (m.size() > model_cap) return fail(NDP_E_INVALID, "model buffer too small");
            memcpy(model, m.data(), m.size() * sizeof(int32_t));
        }
    }
    return NDP_OK;
}

int ndp_dfs_batch(ndp_frontier* f, int n_gpus, int early_exit, uint8_t* verdict, size_t* winner,
                  int32_t* model, size_t model_cap, size_t* model_len, ndp_stats* stats) {
    if (!f) return fail(NDP_E_INVALID, "frontier is null");
    if (n_gpus <= 1)
        return ndp_dfs_shard(f, 0, 1, early_exit, nullptr, verdict, winner, model, model_cap, model_len, nullptr,
                             nullptr, stats);
    int have = 0;
    int rc = ndp_device_count(&have);
    if (rc) return rc;
    if (n_gpus > have) return fail(NDP_E_INVALID, "n_gpus exceeds the devices visible to this process");
    if (n_gpus > 8) return fail(NDP_E_UNSUPPORTED, "at most 8 devices per process");
    if (stats) memset(stats, 0, sizeof *stats);
    if (model_len) *model_len = 0;
    if (winner) *winner = SIZE_MAX;
    // devices: the frontier's home device first, then the others in ordinal order
    std::vector<int> devs{f->device};
    for (int d = 0; d < have && (int)devs.size() < n_gpus; ++d)
        if (d != f->device) devs.push_back(d);
    // peer-visible early-exit words: one per device, every kernel atomically lowers all of them
    std::vector<unsigned long long*> best(n_gpus, nullptr);
    const unsigned long long none = ~0ull;
    for (int i = 0; i < n_gpus; ++i) {
        CUDA_TRY(cudaSetDevice(devs[i]));
        for (int j = 0; j < n_gpus; ++j)
            if (i != j) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
                if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0); if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); }
            }
        CUDA_TRY(cudaMalloc(&best[i], 8));
        CUDA_TRY(cudaMemcpy(best[i], &none, 8, cudaMemcpyHostToDevice));
    }
    // build replicas serially (peer copies from the home device), then search concurrently
    for (int i = 0; i < n_gpus; ++i) {
        ndp_frontier::Replica* rep = nullptr;
        if ((rc = build_replica(f, devs[i], (size_t)i, (size_t)n_gpus, &rep))) return rc;
    }
    std::vector<int> rcs(n_gpus, 0);
    std::vector<std::string> errs(n_gpus);
    std::vector<size_t> wins(n_gpus, SIZE_MAX);
    std::vector<std::vector<int32_t>> models(n_gpus);
    std::vector<ndp_stats> st(n_gpus);
    std::vector<std::thread> th;
    for (int i = 0; i < n_gpus; ++i)
        th.emplace_back([&, i]() {
            memset(&st[i], 0, sizeof(ndp_stats));
            rcs[i] = dfs_on_device(f, devs[i], (size_t)i, (size_t)n_gpus, early_exit, best[i], best.data(), n_gpus,
                                   none, verdict, &wins[i], &models[i], nullptr, nullptr, &st[i]);
            if (rcs[i]) errs[i] = g_err;
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < n_gpus; ++i) { cudaSetDevice(devs[i]); cudaFree(best[i]); }
    cudaSetDevice(f->device);
    for (int i = 0; i < n_gpus; ++i)
        if (rcs[i]) return fail(rcs[i], errs[i]);
    size_t win = SIZE_MAX;
    int who = -1;
    for (int i = 0; i < n_gpus; ++i)
        if (wins[i] < win) { win = wins[i]; who = i; }
    if (stats)
        for (int i = 0; i < n_gpus; ++i) {
            stats->nodes += st[i].nodes; stats->clause_evals += st[i].clause_evals; stats->rebuilds += st[i].rebuilds;
            stats->kernel_ms = std::max(stats->kernel_ms, st[i].kernel_ms); stats->launches += st[i].launches;
            stats->h2d_bytes += st[i].h2d_bytes; stats->d2h_bytes += st[i].d2h_bytes;
        }
    if (winner) *winner = win;
    if (who >= 0) {
        if (model_len) *model_len = models[who].size();
        if (model) {
            if (models[who].size() > model_cap) return fail(NDP_E_INVALID, "model buffer too small");
            memcpy(model, models[who].data(), models[who].size() * sizeof(int32_t));
        }
    }
    return NDP_OK;
}

}  // extern "C"
