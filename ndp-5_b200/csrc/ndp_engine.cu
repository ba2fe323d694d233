// ndp_engine.cu -- host side of the C ABI (include/ndp_b200.h): device arenas, the BFS level
// driver with the reference's exact -d/-q stop rules, DFS launch, result gather.
// Two complete pipelines share the driver (ndp_set_engine / env NDP_ENGINE):
//   scan  : nodes are clause sets (8-byte records) in fixed-stride slots; kernels in ndp_kernels.cuh
//   index : nodes are state slots (2 bit-planes + assignment table) over one occurrence index of the
//           root; kernels in ndp_index.cuh; clause lists are materialised on demand
// There is no CPU fallback in this file: every compute entry point needs a CUDA device.
#include "../../include/ndp_b200.h"
#include "ndp_kernels.cuh"
#include "ndp_index.cuh"
#include "ndp_index_group.cuh"
#include "ndp_split.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace ndp;

// ------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;
thread_local int g_device = 0;
int g_engine = NDP_ENGINE_AUTO;   // process-wide; ndp_set_engine / env NDP_ENGINE
long long g_split = -1;           // process-wide; ndp_set_split / env NDP_SPLIT (-1 auto, 0 off, > 0 piece target)

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

// Device memory comes from the stream-ordered pool of the current device (all work of this
// library runs on the legacy default stream): repeated BFS/DFS calls stop paying cudaMalloc /
// cudaFree (the latter synchronises the device) for their arenas and per-call buffers.
template <class T>
cudaError_t dev_alloc(T** p, size_t bytes) {
    static thread_local int tuned_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (tuned_device != dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        tuned_device = dev;
    }
    return cudaMallocAsync((void**)p, bytes ? bytes : 1, 0);
}
template <class T>
cudaError_t dev_free(T* p) { return p ? cudaFreeAsync((void*)p, 0) : cudaSuccess; }

// Mapped pinned scratch records (BFS level totals): cudaHostAlloc costs about a
// millisecond, so each (host thread, device, slot) gets one allocation that is kept for the
// life of the process.
void* pinned_scratch(int slot, size_t bytes) {
    struct Item { int device, slot; size_t bytes; void* p; };
    static thread_local std::vector<Item> items;
    int dev = 0;
    cudaGetDevice(&dev);
    for (const Item& it : items)
        if (it.device == dev && it.slot == slot && it.bytes >= bytes) return it.p;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    items.push_back(Item{dev, slot, bytes, p});
    return p;
}

// Progress of the DFS calls in flight, one record per device: what the reference's 1 Hz status line shows
// (NDP:1271-1288: remaining queue size, active workers).  `pulled` is a mapped pinned word the DFS kernel
// stores the last unit index into (one posted write per unit); any host thread may read it.
struct ProgressSlot { volatile unsigned* pulled = nullptr; unsigned* d_pulled = nullptr; std::atomic<unsigned long long> total{0}; std::atomic<int> active{0}; };
ProgressSlot g_progress[16];
std::mutex g_progress_mtx;
int progress_arm(int device, size_t total, unsigned** d_word) {
    *d_word = nullptr;
    if (device < 0 || device >= 16) return NDP_OK;
    ProgressSlot& ps = g_progress[device];
    {
        std::lock_guard<std::mutex> lock(g_progress_mtx);
        if (!ps.pulled) {
            void* h = nullptr;
            if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return NDP_OK; }
            ps.pulled = static_cast<volatile unsigned*>(h);
            if (cudaHostGetDevicePointer((void**)&ps.d_pulled, h, 0) != cudaSuccess) { cudaGetLastError(); ps.d_pulled = nullptr; }
        }
    }
    *ps.pulled = 0u;
    ps.total.store(total);
    ps.active.store(1);
    *d_word = ps.d_pulled;
    return NDP_OK;
}
void progress_disarm(int device) {
    if (device >= 0 && device < 16) g_progress[device].active.store(0);
}

int requested_engine() {
    int e = g_engine;
    if (e == NDP_ENGINE_AUTO) {
        const char* s = getenv("NDP_ENGINE");
        if (s && !strcmp(s, "scan")) e = NDP_ENGINE_SCAN;
        else if (s && !strcmp(s, "index")) e = NDP_ENGINE_INDEX;
    }
    return e;
}

struct Entry {        // one frontier element
    int buf;          // which of the two arena buffers holds its slot
    uint32_t slot;
    int32_t n;        // clauses in its set
    int32_t tree;     // index into the choice tree (BFS frontiers) or -1
    int32_t kind, lit;   // what choice() returns for it (state frontiers)
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

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
bool tracing() { static const bool t = getenv("NDP_TRACE") != nullptr; return t; }

struct DevBuf {  // tiny RAII for per-call device allocations (movable, not copyable)
    void* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p) { o.p = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { if (p) dev_free(p); p = o.p; o.p = nullptr; }
        return *this;
    }
    ~DevBuf() { if (p) dev_free(p); }
    template <class T> T* as() { return static_cast<T*>(p); }
};

}  // namespace

struct ndp_frontier {
    int device = 0;
    int max_var = 0;
    std::vector<Entry> entries;
    // ---- clause-slot arena (scan pipeline; also the target of bulk materialisation)
    pclause* buf[2] = {nullptr, nullptr};
    size_t cap_slots[2] = {0, 0};
    size_t stride = 0;                      // clauses per slot
    // ---- state-slot arena (index pipeline)
    bool states = false;                    // entries reference sbuf slots
    unsigned* sbuf[2] = {nullptr, nullptr};
    size_t scap_slots[2] = {0, 0};
    bool materialised = false;              // buf[0] slot k holds entry k's clause list
    // ---- choices: either a parent-pointer tree (BFS) or flat arrays (resume)
    std::vector<int32_t> tree_parent, tree_lit;
    std::vector<int32_t> flat_choices;
    std::vector<size_t> flat_off;
    int32_t *d_tree_parent = nullptr, *d_tree_lit = nullptr;   // device copy of the tree (on `device`)
    size_t d_tree_count = 0;
    bool host_tree_ready = true;            // false while the tree lives only on the device
    // ---- variable renumbering at the ABI (only when the caller's ids do not fit the 16-bit literal codes):
    // the kernels see dense ids 1..V' in first-seen order, every literal that crosses the ABI is mapped back.
    // choice() / resolution only ever look at |l| and sign(l) (NDP:654-781), so any bijection is exact.
    std::vector<int32_t> ext_of;                       // dense variable -> caller's id ([0] unused); empty: identity
    std::unordered_map<int32_t, int32_t> dense_of;     // caller's id -> dense variable
    // ---- root (BFS-born frontiers): what the index engine is built over
    bool has_root = false;
    std::vector<int32_t> root_c3;
    struct RootIndex {                      // per device
        int device = -1;
        pclause* d_root = nullptr;
        uint2* d_hdr = nullptr;
        uint4* d_rows = nullptr;
        unsigned *d_occ_off = nullptr, *d_occ = nullptr;
        IndexRoot ix;
    };
    std::vector<RootIndex> root_indexes;
    size_t root_borrowed = 0;               // the first root_borrowed entries belong to another frontier (piece frontiers)
    // ---- stats
    uint64_t bfs_evals = 0, bfs_launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    double bfs_kernel_ms = 0;
    // ---- scratch for ndp_frontier_get
    std::vector<int32_t> get_clauses, get_choices;
    std::vector<pclause> get_packed;
    pclause* d_get = nullptr;
    const unsigned** d_get_ptr = nullptr;
    // ---- per-(device, shard) DFS descriptors, built lazily
    struct Replica {                        // scan DFS
        int device = -1;
        size_t first = 0, stride = 1, n_local = 0;
        pclause* shard_buf = nullptr;       // private copy of the shard's slots (other devices only)
        const pclause** d_base = nullptr;   // [n_local]
        int32_t* d_n = nullptr;             // [n_local]
        int n_max = 0;
    };
    std::vector<Replica> replicas;
    struct IndexReplica {                   // index DFS
        int device = -1;
        size_t first = 0, stride = 1, n_local = 0;
        unsigned* own_slots = nullptr;      // state slots built or copied for this shard (else BFS slots are used)
        const unsigned** d_slot_ptr = nullptr;   // [n_local]
        BfsNode* d_meta = nullptr;          // [n_local]
        double mean_alive = 0;              // average live clauses per subproblem of the shard
    };
    std::vector<IndexReplica> index_replicas;

    ndp_frontier() {   // pointers into these vectors are handed out: never reallocate them
        root_indexes.reserve(16);
        replicas.reserve(128);
        index_replicas.reserve(128);
    }

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
    ~ndp_frontier() {
        for (auto& r : replicas) {
            cudaSetDevice(r.device);
            if (r.shard_buf) dev_free(r.shard_buf);
            if (r.d_base) dev_free((void*)r.d_base);
            if (r.d_n) dev_free(r.d_n);
        }
        for (auto& r : index_replicas) {
            cudaSetDevice(r.device);
            if (r.own_slots) dev_free(r.own_slots);
            if (r.d_slot_ptr) dev_free((void*)r.d_slot_ptr);
            if (r.d_meta) dev_free(r.d_meta);
        }
        for (size_t i = root_borrowed; i < root_indexes.size(); ++i) {
            RootIndex& r = root_indexes[i];
            cudaSetDevice(r.device);
            if (r.d_root) dev_free(r.d_root);
            if (r.d_hdr) dev_free(r.d_hdr);
            if (r.d_rows) dev_free(r.d_rows);
            if (r.d_occ_off) dev_free(r.d_occ_off);
            if (r.d_occ) dev_free(r.d_occ);
        }
        cudaSetDevice(device);
        if (d_tree_parent) dev_free(d_tree_parent);
        if (d_tree_lit) dev_free(d_tree_lit);
        if (d_get) dev_free(d_get);
        if (d_get_ptr) dev_free((void*)d_get_ptr);
        for (int b = 0; b < 2; ++b) {
            if (buf[b]) dev_free(buf[b]);
            if (sbuf[b]) dev_free(sbuf[b]);
        }
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

// (re)allocate a DEAD arena buffer so that it holds `slots` slots of `slot_bytes`
template <class T>
int ensure_arena(T*& p, size_t& cap, size_t slots, size_t slot_bytes) {
    if (cap >= slots) return NDP_OK;
    if (p) { CUDA_TRY(dev_free(p)); p = nullptr; cap = 0; }
    const size_t want = std::max(slots, (size_t)64);
    CUDA_TRY(dev_alloc(&p, want * slot_bytes));
    cap = want;
    return NDP_OK;
}

// grow a LIVE arena buffer (contents are kept)
template <class T>
int grow_arena_keep(T*& p, size_t& cap, size_t slots, size_t slot_bytes) {
    if (cap >= slots) return NDP_OK;
    T* q = nullptr;
    CUDA_TRY(dev_alloc(&q, slots * slot_bytes));
    if (p && cap) CUDA_TRY(cudaMemcpy(q, p, cap * slot_bytes, cudaMemcpyDeviceToDevice));
    if (p) dev_free(p);
    p = q;
    cap = slots;
    return NDP_OK;
}

// pack an int32 n x 3 clause array; returns false if a literal does not fit the 16-bit codes
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

// ---- variable renumbering (see ndp_frontier::ext_of)
bool needs_renumbering(const int32_t* lits, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (lits[i] > kMaxVar || lits[i] < -kMaxVar) return true;
    return false;
}
// caller's literals -> dense literals.  Unknown variables get the next dense id (first-seen order).
int map_in(ndp_frontier* f, const int32_t* lits, size_t n, int32_t* out) {
    if (f->ext_of.empty()) f->ext_of.push_back(0);
    for (size_t i = 0; i < n; ++i) {
        const int32_t l = lits[i];
        if (l == 0) { out[i] = 0; continue; }
        if (l == INT32_MIN) return fail(NDP_E_UNSUPPORTED, "literal INT32_MIN has no negation");
        const int32_t v = l < 0 ? -l : l;
        auto it = f->dense_of.find(v);
        int32_t d;
        if (it != f->dense_of.end()) d = it->second;
        else {
            if (f->ext_of.size() > (size_t)kMaxVar)
                return fail(NDP_E_UNSUPPORTED, "more than 32766 distinct variables (16-bit literal codes; state slots must fit shared memory)");
            d = (int32_t)f->ext_of.size();
            f->ext_of.push_back(v);
            f->dense_of.emplace(v, d);
        }
        out[i] = l < 0 ? -d : d;
    }
    return NDP_OK;
}
// dense literals -> caller's literals, in place
void map_out(const ndp_frontier* f, int32_t* lits, size_t n) {
    if (f->ext_of.empty()) return;
    for (size_t i = 0; i < n; ++i) {
        const int32_t l = lits[i];
        if (l == 0) continue;
        const int32_t e = f->ext_of[(size_t)(l < 0 ? -l : l)];
        lits[i] = l < 0 ? -e : e;
    }
}

// Root index of a BFS-born frontier on `device`: packed root + merged occurrence lists.
int get_root_index(ndp_frontier* f, int device, const IndexRoot** out) {
    for (auto& r : f->root_indexes)
        if (r.device == device) { *out = &r.ix; return NDP_OK; }
    if (!f->has_root) return fail(NDP_E_INVALID, "this frontier has no root clause array (built by ndp_frontier_from_host)");
    ndp_frontier::RootIndex r;
    r.device = device;
    const size_t n_root = f->root_c3.size() / 3;
    const int V = f->max_var;
    if (n_root >= (1u << 24)) return fail(NDP_E_UNSUPPORTED, "more than 2^24 root clauses");
    std::vector<pclause> root(std::max(n_root, (size_t)1));
    // word-merged rows for clauses that hold the variable once; per-clause entries (occ) for the rest
    std::vector<unsigned> off((size_t)V + 2, 0), occ;
    std::vector<uint2> hdr((size_t)V + 1, make_uint2(0u, 0u));
    std::vector<uint4> rows;
    r.ix.n_zero3 = 0;
    r.ix.first_zero3 = -1;
    {
        struct Occ { unsigned c, n_pos, n_neg; };
        std::vector<unsigned> cnt((size_t)V + 1, 0);
        auto occurrence = [&](size_t c, int j, Occ& o) -> bool {   // one per (clause, variable): at the variable's first position
            const int32_t* l = &f->root_c3[3 * c];
            if (!l[j]) return false;
            for (int i = 0; i < j; ++i)
                if (l[i] && std::abs(l[i]) == std::abs(l[j])) return false;
            o.c = (unsigned)c; o.n_pos = o.n_neg = 0;
            for (int i = 0; i < 3; ++i) {
                if (l[i] == std::abs(l[j])) ++o.n_pos;
                if (l[i] == -std::abs(l[j])) ++o.n_neg;
            }
            return true;
        };
        Occ o;
        for (size_t c = 0; c < n_root; ++c) {
            const int32_t* l = &f->root_c3[3 * c];
            root[c] = pack_clause(l[0], l[1], l[2]);
            if (!l[0] && !l[1] && !l[2]) { ++r.ix.n_zero3; if (r.ix.first_zero3 < 0) r.ix.first_zero3 = (int)c; }
            for (int j = 0; j < 3; ++j)
                if (occurrence(c, j, o)) {
                    const size_t v = (size_t)std::abs(l[j]);
                    if (o.n_pos + o.n_neg == 1) ++cnt[v];        // upper bound of the row length (before merging)
                    else ++off[v + 1];
                }
        }
        // counting sort by variable; within a variable the entries stay in clause order, so the
        // clauses of one plane word are adjacent and merge into one row entry
        std::vector<size_t> start((size_t)V + 2, 0);
        for (int v = 0; v <= V; ++v) { start[(size_t)v + 1] = start[v] + cnt[v]; off[(size_t)v + 1] += off[v]; }
        std::vector<uint4> raw(start[(size_t)V + 1]);
        std::vector<size_t> fill(start.begin(), start.end() - 1);
        occ.resize(off[(size_t)V + 1]);
        std::vector<unsigned> cursor(off.begin(), off.end() - 1);
        for (size_t c = 0; c < n_root; ++c) {
            const int32_t* l = &f->root_c3[3 * c];
            for (int j = 0; j < 3; ++j)
                if (occurrence(c, j, o)) {
                    const size_t v = (size_t)std::abs(l[j]);
                    const unsigned bit = 1u << (c & 31);
                    if (o.n_pos + o.n_neg == 1) raw[fill[v]++] = make_uint4((unsigned)(c >> 5), o.n_pos ? bit : 0u, o.n_neg ? bit : 0u, 0u);
                    else occ[cursor[v]++] = o.c | (o.n_pos << 24) | (o.n_neg << 26);
                }
        }
        rows.reserve(raw.size());
        for (int v = 0; v <= V; ++v) {
            const size_t first = rows.size();
            for (size_t i = start[v]; i < start[(size_t)v + 1]; ++i) {
                if (rows.size() > first && rows.back().x == raw[i].x) { rows.back().y |= raw[i].y; rows.back().z |= raw[i].z; }
                else rows.push_back(raw[i]);
            }
            const bool multi = off[(size_t)v + 1] > off[v];
            hdr[v] = make_uint2((unsigned)first, (unsigned)(rows.size() - first) | (multi ? 0x80000000u : 0u));
        }
        if (rows.empty()) rows.push_back(make_uint4(0u, 0u, 0u, 0u));
    }
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(dev_alloc(&r.d_root, root.size() * sizeof(pclause)));
    CUDA_TRY(dev_alloc(&r.d_hdr, hdr.size() * sizeof(uint2)));
    CUDA_TRY(dev_alloc(&r.d_rows, rows.size() * sizeof(uint4)));
    CUDA_TRY(dev_alloc(&r.d_occ_off, off.size() * 4));
    CUDA_TRY(dev_alloc(&r.d_occ, std::max(occ.size(), (size_t)1) * 4));
    CUDA_TRY(cudaMemcpy(r.d_root, root.data(), root.size() * sizeof(pclause), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_hdr, hdr.data(), hdr.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_rows, rows.data(), rows.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_occ_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    if (!occ.empty()) CUDA_TRY(cudaMemcpy(r.d_occ, occ.data(), occ.size() * 4, cudaMemcpyHostToDevice));
    f->h2d_bytes += root.size() * sizeof(pclause) + hdr.size() * sizeof(uint2) + rows.size() * sizeof(uint4) + off.size() * 4 +
                    occ.size() * 4;
    r.ix.row_hdr = r.d_hdr;
    r.ix.rows = r.d_rows;
    r.ix.root = r.d_root;
    r.ix.n_root = (int)n_root;
    r.ix.occ_off = r.d_occ_off;
    r.ix.occ = r.d_occ;
    r.ix.bm_words = (int)((n_root + 31) / 32);
    r.ix.table_words = (V + 1 + 15) / 16;
    r.ix.sum_words = (r.ix.bm_words + 31) / 32;
    if (r.ix.sum_words > 32 || getenv("NDP_NO_SUMMARY")) r.ix.sum_words = 0;   // > 32768 clauses: plain scan of B[2]
    r.ix.table_off = 2 * r.ix.bm_words;
    r.ix.sum_off = r.ix.table_off + r.ix.table_words;
    r.ix.slot_words = (r.ix.table_off + r.ix.table_words + r.ix.sum_words + 3) / 4 * 4;
    if (r.ix.slot_words == 0) r.ix.slot_words = 4;
    r.ix.plain = r.ix.sum_words > 0 && r.ix.n_zero3 == 0 && occ.empty() && !getenv("NDP_NO_PLAIN");
    f->root_indexes.push_back(r);
    *out = &f->root_indexes.back().ix;
    return NDP_OK;
}

// Device-resident bookkeeping of the BFS level loop (freed on scope exit).
struct BfsLevelState {
    BfsNode* nodes[2] = {nullptr, nullptr};
    int32_t* node_tree[2] = {nullptr, nullptr};
    BfsChild* children = nullptr;
    int32_t *tree_parent = nullptr, *tree_lit = nullptr;
    BfsLevelTotals* totals = nullptr;        // device view of the host-mapped totals record
    volatile BfsLevelTotals* h_totals = nullptr;
    size_t node_cap = 0, tree_cap = 0, tree_count = 0;
    int cur = 0;

    template <class T> static int grow(T*& p, size_t old_n, size_t new_n) {
        T* q = nullptr;
        CUDA_TRY(dev_alloc(&q, new_n * sizeof(T)));
        if (p && old_n) CUDA_TRY(cudaMemcpy(q, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice));
        if (p) dev_free(p);
        p = q;
        return NDP_OK;
    }
    int reserve_nodes(size_t want) {
        if (!totals) {
            // the per-level totals are written straight into mapped pinned host memory: the host
            // only has to wait for the stream, no copy is issued
            void* h = pinned_scratch(0, sizeof(BfsLevelTotals));
            if (!h) return fail(NDP_E_NOMEM, "pinned host scratch");
            h_totals = static_cast<volatile BfsLevelTotals*>(h);
            CUDA_TRY(cudaHostGetDevicePointer((void**)&totals, h, 0));
        }
        if (want <= node_cap) return NDP_OK;
        size_t cap = std::max(want, std::max(node_cap * 2, (size_t)4096));
        int rc;
        for (int b = 0; b < 2; ++b) {
            if ((rc = grow(nodes[b], node_cap, cap))) return rc;
            if ((rc = grow(node_tree[b], node_cap, cap))) return rc;
        }
        if (children) dev_free(children);
        children = nullptr;
        CUDA_TRY(dev_alloc(&children, cap * sizeof(BfsChild)));
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
        for (int b = 0; b < 2; ++b) { if (nodes[b]) dev_free(nodes[b]); if (node_tree[b]) dev_free(node_tree[b]); }
        if (children) dev_free(children);
        if (tree_parent) dev_free(tree_parent);
        if (tree_lit) dev_free(tree_lit);
    }
};

int ensure_host_tree(ndp_frontier* f) {
    if (f->host_tree_ready) return NDP_OK;
    CUDA_TRY(cudaSetDevice(f->device));
    f->tree_parent.resize(f->d_tree_count);
    f->tree_lit.resize(f->d_tree_count);
    CUDA_TRY(cudaMemcpy(f->tree_parent.data(), f->d_tree_parent, f->d_tree_count * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(f->tree_lit.data(), f->d_tree_lit, f->d_tree_count * 4, cudaMemcpyDeviceToHost));
    f->d2h_bytes += f->d_tree_count * 8;
    f->host_tree_ready = true;
    return NDP_OK;
}

// choices of ONE element without bringing the whole tree home: walk it on the device
int frontier_choices(ndp_frontier* f, size_t k, std::vector<int32_t>& out) {
    const Entry& e = f->entries[k];
    if (f->host_tree_ready || e.tree < 0) { f->choices_of(k, out); return NDP_OK; }
    CUDA_TRY(cudaSetDevice(f->device));
    const int cap = std::max(f->max_var, 1) + 8;
    DevBuf d_out, d_len;
    CUDA_TRY(dev_alloc(&d_out.p, (size_t)cap * 4));
    CUDA_TRY(dev_alloc(&d_len.p, 4));
    tree_path_kernel<<<1, 1>>>(f->d_tree_parent, f->d_tree_lit, e.tree, d_out.as<int32_t>(), cap, d_len.as<int32_t>());
    CUDA_TRY(cudaGetLastError());
    int32_t len = 0;
    CUDA_TRY(cudaMemcpy(&len, d_len.p, 4, cudaMemcpyDeviceToHost));
    if (len > cap) {   // deeper than any variable count allows: fall back to the full tree
        int rc = ensure_host_tree(f);
        if (rc) return rc;
        f->choices_of(k, out);
        return NDP_OK;
    }
    out.resize((size_t)len);
    if (len) CUDA_TRY(cudaMemcpy(out.data(), d_out.p, (size_t)len * 4, cudaMemcpyDeviceToHost));
    std::reverse(out.begin(), out.end());
    return NDP_OK;
}

// Clause lists of a state frontier: element k -> out slot k (stride records each).
int materialise(ndp_frontier* f, const std::vector<size_t>& which, pclause* d_out, size_t stride) {
    const IndexRoot* ix = nullptr;
    int rc = get_root_index(f, f->device, &ix);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(f->device));
    std::vector<const unsigned*> ptrs(which.size());
    for (size_t i = 0; i < which.size(); ++i) {
        const Entry& e = f->entries[which[i]];
        ptrs[i] = f->sbuf[e.buf] + (size_t)e.slot * ix->slot_words;
    }
    DevBuf d_ptr;
    CUDA_TRY(dev_alloc(&d_ptr.p, std::max(ptrs.size(), (size_t)1) * sizeof(unsigned*)));
    CUDA_TRY(cudaMemcpy(d_ptr.p, ptrs.data(), ptrs.size() * sizeof(unsigned*), cudaMemcpyHostToDevice));
    const int wpb = 4;
    const int grid = (int)std::min<size_t>((which.size() + wpb - 1) / wpb, (size_t)148 * 16);
    index_materialise_kernel<<<std::max(grid, 1), wpb * 32>>>(*ix, d_ptr.as<const unsigned*>(), (int)which.size(), d_out,
                                                              stride, nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return NDP_OK;
}

// Make buf[0] slot k hold entry k's clause list (what the scan DFS and bulk export read).
int ensure_materialised(ndp_frontier* f) {
    if (!f->states || f->materialised) return NDP_OK;
    int n_max = 1;
    for (const Entry& e : f->entries) n_max = std::max(n_max, e.n);
    f->stride = (size_t)pad_up(n_max);
    CUDA_TRY(cudaSetDevice(f->device));
    int rc = ensure_arena(f->buf[0], f->cap_slots[0], std::max(f->entries.size(), (size_t)1), f->stride * sizeof(pclause));
    if (rc) return rc;
    std::vector<size_t> all(f->entries.size());
    for (size_t k = 0; k < all.size(); ++k) all[k] = k;
    if ((rc = materialise(f, all, f->buf[0], f->stride))) return rc;
    f->materialised = true;
    return NDP_OK;
}

const pclause* clause_slot(const ndp_frontier* f, size_t k) {
    const Entry& e = f->entries[k];
    return f->states ? f->buf[0] + k * f->stride : f->buf[e.buf] + (size_t)e.slot * f->stride;
}

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" {

const char* ndp_last_error(void) { return g_err.c_str(); }
const char* ndp_version(void) { return "ndp-b200 0.2 (NDP 5.6.7 hot path, sm_100a; engines: scan, index)"; }

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
int ndp_set_split(long long pieces) {
    if (pieces < -1) return fail(NDP_E_INVALID, "split target must be -1 (auto), 0 (off) or a piece count");
    g_split = pieces;
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
    const bool use_index = requested_engine() != NDP_ENGINE_SCAN;

    const double t_start = now_ms();
    double t_safe = 0, t_unsafe = 0;
    size_t n_safe = 0, n_unsafe = 0;
    std::unique_ptr<ndp_frontier> f(new ndp_frontier);
    f->device = g_device;
    f->stride = std::max(n_clauses, (size_t)1);
    std::vector<pclause> packed(f->stride);
    std::vector<int32_t> dense;                   // ids beyond the 16-bit codes: renumber (sparse ids are fine, > 32766 distinct are not)
    if (needs_renumbering(clauses3, 3 * n_clauses)) {
        dense.resize(3 * n_clauses);
        if ((rc = map_in(f.get(), clauses3, 3 * n_clauses, dense.data()))) return rc;
        clauses3 = dense.data();
    }
    if (!pack_host(clauses3, n_clauses, packed.data(), f->max_var))
        return fail(NDP_E_UNSUPPORTED, "literal magnitude exceeds 32766 (16-bit clause packing)");
    f->root_c3.assign(clauses3, clauses3 + 3 * n_clauses);
    f->has_root = true;
    f->states = use_index;

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
    const IndexRoot* ix = nullptr;
    size_t slot_bytes = f->stride * sizeof(pclause);
    BfsNode root{0, (int32_t)n_clauses, 0, 0};
    host_setinfo(clauses3, n_clauses, root.kind, root.lit);
    if ((status = st.reserve_nodes(1)) || (status = st.reserve_tree(1))) return status;
    if (use_index) {
        if ((rc = get_root_index(f.get(), f->device, &ix))) return rc;
        slot_bytes = (size_t)ix->slot_words * 4;
        if (8 * slot_bytes > 227 * 1024)
            return fail(NDP_E_UNSUPPORTED, "root too large for the index engine's shared-memory state slots; use the scan engine");
        CUDA_TRY(cudaFuncSetAttribute(bfs_index_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * slot_bytes)));
        CUDA_TRY(cudaFuncSetAttribute(index_states_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * slot_bytes)));
        if ((rc = ensure_arena(f->sbuf[0], f->scap_slots[0], 64, slot_bytes))) return rc;
        // root state = the bitmaps of the empty assignment
        DevBuf d_tab;
        CUDA_TRY(dev_alloc(&d_tab.p, std::max(ix->table_words, 1) * 4));
        CUDA_TRY(cudaMemset(d_tab.p, 0, std::max(ix->table_words, 1) * 4));
        index_states_kernel<<<1, 32, ix->slot_words * 4>>>(*ix, d_tab.as<unsigned>(), f->sbuf[0], 1, st.nodes[0]);
        CUDA_TRY(cudaGetLastError());
        BfsNode dev_root;
        CUDA_TRY(cudaMemcpy(&dev_root, st.nodes[0], sizeof dev_root, cudaMemcpyDeviceToHost));
        if (dev_root.n != root.n || dev_root.kind != root.kind || dev_root.lit != root.lit)
            return fail(NDP_E_CUDA, "index root state disagrees with the host's choice()");
        f->bfs_launches += 1;
    } else {
        if ((rc = ensure_arena(f->buf[0], f->cap_slots[0], 64, slot_bytes))) return rc;
        CUDA_TRY(cudaMemcpy(f->buf[0], packed.data(), n_clauses * sizeof(pclause), cudaMemcpyHostToDevice));
        f->h2d_bytes += n_clauses * sizeof(pclause);
        CUDA_TRY(cudaMemcpy(st.nodes[0], &root, sizeof root, cudaMemcpyHostToDevice));
    }
    {
        int32_t zero = 0, minus1 = -1;
        CUDA_TRY(cudaMemcpy(st.node_tree[0], &zero, 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(st.tree_parent, &minus1, 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(st.tree_lit, &zero, 4, cudaMemcpyHostToDevice));
        st.tree_count = 1;
        f->h2d_bytes += sizeof root + 12;
    }
    // persistent level loop (index engine): cooperative launch, one CTA per SM
    bool persist_ok = false, persist_rearm = false;
    size_t persist_cap = 0;
    DevBuf d_pstate, d_pscratch, d_psnaps, d_ppaths;
    size_t ppaths_cap = 0;
    int window_max = kWindowMax;
    if (const char* e = getenv("NDP_BFS_WINDOW")) window_max = std::max(1, std::min(kWindowMax, atoi(e)));
    int max_block_nodes = kPersistMaxBlockNodes;   // test knob: exercise the wide-level path on small inputs
    if (const char* e = getenv("NDP_PERSIST_BLOCK_NODES")) max_block_nodes = std::max(1, std::min(kPersistMaxBlockNodes, atoi(e)));
    int coop_grid = 0, coop_warps = 0;
    if (use_index && !getenv("NDP_NO_PERSIST")) {
        int coop = 0, sms = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, f->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
        // one CTA per SM, as many warps as state slots fit its shared memory (<= 32)
        coop_warps = (int)std::min<size_t>(32, (200 * 1024) / slot_bytes);
        if (const char* e = getenv("NDP_PERSIST_WARPS")) coop_warps = std::max(1, std::min(coop_warps, atoi(e)));
        if (coop && coop_warps >= 4) {
            const void* pkern = ix->plain ? (const void*)bfs_index_persistent_kernel<true> : (const void*)bfs_index_persistent_kernel<false>;
            CUDA_TRY(cudaFuncSetAttribute(pkern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(coop_warps * slot_bytes)));
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pkern, coop_warps * 32,
                                                                   coop_warps * slot_bytes));
            if (per_sm >= 1) {
                coop_grid = std::min(sms, 160);   // PersistScratch holds 160 CTA records
                CUDA_TRY(dev_alloc(&d_pstate.p, sizeof(PersistState)));
                CUDA_TRY(dev_alloc(&d_pscratch.p, sizeof(PersistScratch)));
                if (window_max > 1)   // snapshots of pending RA siblings, one stack per resident warp
                    CUDA_TRY(dev_alloc(&d_psnaps.p, (size_t)coop_grid * coop_warps * kWindowPend * slot_bytes));
                persist_ok = true;
                const size_t budget = ((size_t)4 << 30) / slot_bytes;   // <= 4 GB per buffer up front
                persist_cap = Q > 0 ? std::min<size_t>(2 * (size_t)Q, budget) : std::min<size_t>(8192, budget);
                persist_cap = std::max<size_t>(persist_cap, 64);
            }
        }
    }
    cudaEvent_t ev_loop0, ev_loop1;        // device timeline of the whole level loop
    CUDA_TRY(cudaEventCreate(&ev_loop0));
    CUDA_TRY(cudaEventCreate(&ev_loop1));
    CUDA_TRY(cudaEventRecord(ev_loop0, 0));

    std::vector<BfsNode> h_nodes;          // current level, only materialised on the host when needed
    std::vector<int32_t> h_node_tree;
    std::vector<BfsChild> h_children;
    std::vector<Entry> final_entries;
    struct Pushed { int32_t parent_tree, lit; };
    std::vector<Pushed> host_tree_tail;    // tree entries created by the host walk of the stop level

    // Level-synchronous restatement of the FIFO loop NDP:893-937: a FIFO
    // queue only ever holds the unexpanded rest of level L followed by the level-L+1 children
    // pushed so far, so walking one level in order with the running queue size `s` reproduces
    // every stop test at the exact pop it would fire.
    const double t_loop = now_ms();
    for (;;) {
        const double t_level = now_ms();
        if (m == 0) break;                                           // queue empty (NDP:893)
        const bool stop_at_top = (Q > 0 && m >= (size_t)Q) || (D == -1 && !ovr);   // NDP:894-897 before the first pop
        size_t need = m;
        if (!stop_at_top && Q <= 0) {                                // -d budget known in advance
            long long R = (long long)D - it;
            need = (size_t)std::min<long long>((long long)m, std::max<long long>(R, 1));
        }
        // can any stop test fire inside this level?  -q: the queue never exceeds m + pushes <= 2m;
        // -d: the level is consumed whole iff it + m < D.
        const bool safe = !stop_at_top && (Q > 0 ? 2 * m < (size_t)Q : (it + (long long)m < (long long)D));
        if (persist_ok && !stop_at_top && (Q > 0 || it + (long long)m < (long long)D)) {
            // hand the level loop to the device until a stop level, an empty queue or a full arena
            if (2 * m > persist_cap) persist_cap = std::max(persist_cap * 2, 4 * m);
            if ((status = grow_arena_keep(f->sbuf[cur_buf], f->scap_slots[cur_buf], persist_cap, slot_bytes))) break;
            status = ensure_arena(f->sbuf[cur_buf ^ 1], f->scap_slots[cur_buf ^ 1], persist_cap, slot_bytes);
            if (status == NDP_E_NOMEM) { cudaGetLastError(); g_err.clear(); status = NDP_OK; persist_ok = false; }
            if (status) break;
            if (persist_ok) {
                const size_t tree_room = std::max<size_t>(16 * persist_cap, (size_t)5 << 20);   // a window of K levels appends K entries per output
                if ((status = st.reserve_nodes(persist_cap)) || (status = st.reserve_tree(st.tree_count + tree_room))) break;
                PersistState ps;
                memset(&ps, 0, sizeof ps);
                ps.m = (int)m; ps.cur = st.cur; ps.cur_buf = cur_buf;
                ps.it = it; ps.task_count = task_count; ps.tree_count = st.tree_count;
                ps.cap_slots = std::min(f->scap_slots[0], f->scap_slots[1]);
                ps.cap_nodes = st.node_cap;
                ps.cap_tree = st.tree_cap;
                CUDA_TRY(cudaMemcpyAsync(d_pstate.p, &ps, sizeof ps, cudaMemcpyHostToDevice, 0));
                CUDA_TRY(cudaMemsetAsync(d_pscratch.p, 0, sizeof(PersistScratch), 0));
                IndexRoot ixv = *ix;
                unsigned *a0 = f->sbuf[0], *a1 = f->sbuf[1];
                BfsNode *n0 = st.nodes[0], *n1 = st.nodes[1];
                int32_t *t0 = st.node_tree[0], *t1 = st.node_tree[1], *tp = st.tree_parent, *tl = st.tree_lit;
                BfsChild* chd = st.children;
                int Dv = D, Qv = Q, ovrv = ovr ? 1 : 0, maxl = 1 << 20;
                PersistState* dps = d_pstate.as<PersistState>();
                PersistScratch* dsc = d_pscratch.as<PersistScratch>();
                if (window_max > 1 && ppaths_cap < st.node_cap) {   // K choices per output of a window
                    if (d_ppaths.p) { dev_free(d_ppaths.p); d_ppaths.p = nullptr; }
                    CUDA_TRY(dev_alloc(&d_ppaths.p, st.node_cap * kWindowMax * sizeof(short)));
                    ppaths_cap = st.node_cap;
                }
                unsigned* snaps = d_psnaps.as<unsigned>();
                short* ppaths = d_ppaths.as<short>();
                void* args[] = {&ixv, &a0, &a1, &n0, &n1, &t0, &t1, &chd, &tp, &tl, &Dv, &Qv, &ovrv, &maxl, &dps, &dsc,
                                &snaps, &ppaths, &window_max, &max_block_nodes};
                cudaError_t le = cudaLaunchCooperativeKernel(ix->plain ? (const void*)bfs_index_persistent_kernel<true>
                                                                       : (const void*)bfs_index_persistent_kernel<false>, dim3(coop_grid),
                                                             dim3(coop_warps * 32), args, coop_warps * slot_bytes, 0);
                if (le == cudaSuccess) le = cudaStreamSynchronize(0);
                if (le != cudaSuccess) { status = fail(NDP_E_CUDA, std::string("BFS persistent kernel: ") + cudaGetErrorString(le)); break; }
                CUDA_TRY(cudaMemcpy(&ps, d_pstate.p, sizeof ps, cudaMemcpyDeviceToHost));
                f->bfs_launches += 1;
                f->h2d_bytes += sizeof ps;
                f->d2h_bytes += sizeof ps;
                if (tracing())
                    fprintf(stderr, "[ndp_bfs] persistent: %d levels, m %zu -> %d, exit %d, windows ok %llu redone %llu (last K %d) | "
                                    "CTA 0: expand %.2f, scan+look-back %.2f, emit %.2f, grid.sync %.2f, commit %.2f ms\n",
                            ps.levels_done, m, ps.m, ps.exit_reason, ps.windows_ok, ps.windows_redone, ps.window,
                            ps.t_phase[0] / 1e6, ps.t_phase[1] / 1e6, ps.t_phase[2] / 1e6, ps.t_phase[3] / 1e6, ps.t_phase[4] / 1e6);
                m = (size_t)ps.m; st.cur = ps.cur; cur_buf = ps.cur_buf;
                it = ps.it; task_count = ps.task_count; f->bfs_evals += ps.evals; st.tree_count = ps.tree_count;
                n_safe += (size_t)ps.levels_done;
                t_safe += now_ms() - t_level;
                if (ps.exit_reason == PEXIT_CAPACITY) {
                    // the choice tree is simply re-reserved on re-entry; slots / node lists grow here
                    if (2 * m > persist_cap) {
                        const size_t bound = Q > 0 ? 2 * (size_t)Q : SIZE_MAX;
                        persist_cap = std::min(std::max(persist_cap * 2, 4 * m), std::max(bound, 2 * m));
                    }
                    continue;
                }
                if (ps.exit_reason == PEXIT_EMPTY || ps.exit_reason == PEXIT_LEVELS) continue;
                // PEXIT_HOST_LEVEL (a stop test can fire) / PEXIT_WIDE (more nodes per CTA than the kernel's
                // shared scan arrays hold): this level goes through the per-level path, then the device
                // loop is re-armed
                persist_ok = false;
                persist_rearm = true;
                continue;
            }
        }
        if (stop_at_top) {
            if ((status = st.download_level(m, h_nodes, h_node_tree, f->d2h_bytes))) break;
            for (size_t j = 0; j < m; ++j)
                final_entries.push_back(Entry{cur_buf, h_nodes[j].slot, h_nodes[j].n, h_node_tree[j], h_nodes[j].kind, h_nodes[j].lit});
            break;
        }
        {
            // grow the (dead) target buffer geometrically, but never past what the stop rule
            // allows a level to reach: width <= Q under -q, <= D+1 under -d
            size_t want = 2 * need;
            size_t& cap = use_index ? f->scap_slots[cur_buf ^ 1] : f->cap_slots[cur_buf ^ 1];
            if (cap < want) {
                size_t bound = SIZE_MAX;
                if (Q > 0) bound = 2 * (size_t)Q;
                else if (D > 0) bound = 2 * ((size_t)D + 1);
                want = std::min(std::max(want, cap * 2), std::max(bound, want));
            }
            auto grow = [&](size_t slots) {
                return use_index ? ensure_arena(f->sbuf[cur_buf ^ 1], f->scap_slots[cur_buf ^ 1], slots, slot_bytes)
                                 : ensure_arena(f->buf[cur_buf ^ 1], f->cap_slots[cur_buf ^ 1], slots, slot_bytes);
            };
            status = grow(want);
            if (status == NDP_E_NOMEM && want > 2 * need) {   // head-room did not fit: take the exact size
                cudaGetLastError();
                status = grow(2 * need);
            }
            if (status) break;
        }
        if ((status = st.reserve_nodes(2 * need)) || (status = st.reserve_tree(st.tree_count + 2 * need))) break;
        const int warps_per_block = 8;
        int grid = (int)std::min<size_t>((need + warps_per_block - 1) / warps_per_block, (size_t)148 * 16);
        if (use_index)
            bfs_index_expand_kernel<<<grid, warps_per_block * 32, (size_t)warps_per_block * ix->slot_words * 4>>>(
                *ix, f->sbuf[cur_buf], f->sbuf[cur_buf ^ 1], st.nodes[st.cur], (int)need, st.children);
        else
            bfs_expand_kernel<<<grid, warps_per_block * 32>>>(f->buf[cur_buf], f->buf[cur_buf ^ 1], f->stride,
                                                              st.nodes[st.cur], (int)need, st.children);
        bfs_compact_kernel<<<1, 1024>>>(st.nodes[st.cur], st.node_tree[st.cur], st.children, (int)need,
                                        st.nodes[st.cur ^ 1], st.node_tree[st.cur ^ 1], st.tree_parent, st.tree_lit,
                                        (int)st.tree_count, Q, st.totals);
        f->bfs_launches += 2;
        cudaError_t ce = cudaStreamSynchronize(0);
        if (ce != cudaSuccess) { status = fail(NDP_E_CUDA, std::string("BFS level: ") + cudaGetErrorString(ce)); break; }
        BfsLevelTotals tot;
        memcpy(&tot, (const void*)st.h_totals, sizeof tot);
        f->d2h_bytes += sizeof tot;
        // -q: the device checked the "queue size >= Q before a pop" test at every node of the level
        if (safe || (Q > 0 && !tot.stop_q)) {
            // every node of the level is expanded; only the totals matter
            it += (long long)tot.expanded;
            task_count += (long long)tot.m_next;
            f->bfs_evals += tot.evals;
            st.tree_count += tot.m_next;
            st.cur ^= 1;
            cur_buf ^= 1;
            m = tot.m_next;
            t_safe += now_ms() - t_level;
            ++n_safe;
            if (persist_rearm) { persist_ok = true; persist_rearm = false; }
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
                next_entries.push_back(Entry{cur_buf ^ 1, (uint32_t)(2 * j + b), c.n, tree, c.kind, c.lit});
                ++task_count;
                ++s;
            }
            ++it;                                                    // NDP:926
            if (Q <= 0 && it >= D) { stopped = true; break; }        // NDP:935-936
        }
        if (stopped || E < m) {
            // frontier = unexpanded rest of the current level, then the children pushed so far
            for (size_t j = E; j < m; ++j)
                final_entries.push_back(Entry{cur_buf, h_nodes[j].slot, h_nodes[j].n, h_node_tree[j], h_nodes[j].kind, h_nodes[j].lit});
            final_entries.insert(final_entries.end(), next_entries.begin(), next_entries.end());
            break;
        }
        // no stop fired: the device-side compaction of this level is exactly what the walk produced
        host_tree_tail.clear();
        st.tree_count += tot.m_next;
        st.cur ^= 1;
        cur_buf ^= 1;
        m = tot.m_next;
        t_unsafe += now_ms() - t_level;
        ++n_unsafe;
    }
    cudaEventRecord(ev_loop1, 0);
    cudaDeviceSynchronize();
    const double t_loop_end = now_ms();
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev_loop0, ev_loop1) == cudaSuccess) f->bfs_kernel_ms = ms;
        cudaEventDestroy(ev_loop0);
        cudaEventDestroy(ev_loop1);
    }
    if (status) return status;

    // The choice tree stays on the device (completed with the stop level's tail); the host copy is
    // fetched lazily, only if choices of many elements are asked for (ndp_frontier_get / -s).
    f->entries.swap(final_entries);
    if (!host_tree_tail.empty()) {
        if ((status = st.reserve_tree(st.tree_count + host_tree_tail.size()))) return status;
        std::vector<int32_t> tp(host_tree_tail.size()), tl(host_tree_tail.size());
        for (size_t i = 0; i < host_tree_tail.size(); ++i) { tp[i] = host_tree_tail[i].parent_tree; tl[i] = host_tree_tail[i].lit; }
        CUDA_TRY(cudaMemcpy(st.tree_parent + st.tree_count, tp.data(), tp.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(st.tree_lit + st.tree_count, tl.data(), tl.size() * 4, cudaMemcpyHostToDevice));
        f->h2d_bytes += host_tree_tail.size() * 8;
    }
    f->d_tree_count = st.tree_count + host_tree_tail.size();
    f->host_tree_ready = false;
    f->d_tree_parent = st.tree_parent;
    f->d_tree_lit = st.tree_lit;
    st.tree_parent = nullptr;   // ownership moves to the frontier
    st.tree_lit = nullptr;

    if (tracing())
        fprintf(stderr, "[ndp_bfs] setup %.2f ms | loop %.2f ms (safe levels %zu: %.2f ms, walked levels %zu: %.2f ms) | "
                        "finish %.2f ms | kernels %.2f ms\n",
                t_loop - t_start, t_loop_end - t_loop, n_safe, t_safe, n_unsafe, t_unsafe, now_ms() - t_loop_end,
                f->bfs_kernel_ms);
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
    if (n_choices) {
        int rc = ensure_host_tree(const_cast<ndp_frontier*>(f));
        if (rc) return rc;
        *n_choices = f->nchoices_of(k);
    }
    return NDP_OK;
}

int ndp_frontier_get(ndp_frontier* f, size_t k, const int32_t** clauses3, size_t* n_clauses,
                     const int32_t** choices, size_t* n_choices) {
    if (!f || k >= f->entries.size()) return fail(NDP_E_INVALID, "frontier index out of range");
    const Entry& e = f->entries[k];
    if (clauses3 || n_clauses) {
        CUDA_TRY(cudaSetDevice(f->device));
        const pclause* src;
        if (f->states && !f->materialised) {
            // one element: filter the root by this node's assignment into a scratch slot
            const size_t cap = (size_t)pad_up((int)(f->root_c3.size() / 3) + 1);
            if (!f->d_get) CUDA_TRY(dev_alloc(&f->d_get, cap * sizeof(pclause)));
            int rc = materialise(f, std::vector<size_t>{k}, f->d_get, cap);
            if (rc) return rc;
            src = f->d_get;
        } else {
            src = clause_slot(f, k);
        }
        f->get_packed.resize((size_t)std::max(e.n, 1));
        CUDA_TRY(cudaMemcpy(f->get_packed.data(), src, (size_t)e.n * sizeof(pclause), cudaMemcpyDeviceToHost));
        f->get_clauses.resize((size_t)e.n * 3 + 1);
        for (int32_t c = 0; c < e.n; ++c) {
            int l0, l1, l2;
            unpack_clause(f->get_packed[c], l0, l1, l2);
            f->get_clauses[3 * c] = l0;
            f->get_clauses[3 * c + 1] = l1;
            f->get_clauses[3 * c + 2] = l2;
        }
        map_out(f, f->get_clauses.data(), (size_t)e.n * 3);
        if (clauses3) *clauses3 = f->get_clauses.data();
        if (n_clauses) *n_clauses = (size_t)e.n;
    }
    if (choices || n_choices) {
        int rc = ensure_host_tree(f);
        if (rc) return rc;
        f->choices_of(k, f->get_choices);
        map_out(f, f->get_choices.data(), f->get_choices.size());
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
    if ((rc = ensure_arena(f->buf[0], f->cap_slots[0], std::max(count, (size_t)1), stride * sizeof(pclause)))) return rc;
    // stage in bounded chunks so a multi-GB resume does not double its footprint on the host
    const size_t chunk_slots = std::max<size_t>(1, ((size_t)64 << 20) / (stride * sizeof(pclause)));
    std::vector<pclause> stage(chunk_slots * stride);
    const bool renumber = count && (needs_renumbering(clauses3 + 3 * clause_offsets[0], 3 * (clause_offsets[count] - clause_offsets[0])) ||
                                    needs_renumbering(choices + choice_offsets[0], choice_offsets[count] - choice_offsets[0]));
    std::vector<int32_t> dense;
    for (size_t k0 = 0; k0 < count; k0 += chunk_slots) {
        size_t k1 = std::min(count, k0 + chunk_slots);
        for (size_t k = k0; k < k1; ++k) {
            size_t n = clause_offsets[k + 1] - clause_offsets[k];
            const int32_t* c3 = clauses3 + 3 * clause_offsets[k];
            if (renumber) {
                dense.resize(3 * n + 1);
                if ((rc = map_in(f.get(), c3, 3 * n, dense.data()))) return rc;
                c3 = dense.data();
            }
            if (!pack_host(c3, n, stage.data() + (k - k0) * stride, f->max_var))
                return fail(NDP_E_UNSUPPORTED, "literal magnitude exceeds 32766 (16-bit clause packing)");
            f->entries.push_back(Entry{0, (uint32_t)k, (int32_t)n, -1, 0, 0});
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
    if (renumber && (rc = map_in(f.get(), f->flat_choices.data(), f->flat_choices.size(), f->flat_choices.data()))) return rc;
    for (int32_t l : f->flat_choices) {
        if (l > kMaxVar || l < -kMaxVar) return fail(NDP_E_UNSUPPORTED, "choice magnitude exceeds 32766");
        f->max_var = std::max(f->max_var, std::abs(l));
    }
    *out = f.release();
    return NDP_OK;
}

int ndp_frontier_attach_root(ndp_frontier* f, const int32_t* clauses3, size_t n_clauses) {
    if (!f) return fail(NDP_E_INVALID, "frontier is null");
    if (f->has_root) return NDP_OK;                                   // BFS-born: nothing to do
    if (!clauses3 && n_clauses) return fail(NDP_E_INVALID, "clauses3 is null");
    if (n_clauses > (size_t)INT_MAX / 4) return fail(NDP_E_UNSUPPORTED, "too many clauses");
    int rc = check_device();
    if (rc) return rc;
    if (f->device != g_device) return fail(NDP_E_INVALID, "frontier lives on another device");
    const size_t count = f->entries.size();
    if (count == 0 || count > (size_t)INT_MAX) return NDP_OK;
    // candidate root: everything below is undone if the frontier turns out not to descend from it
    std::vector<pclause> packed(std::max(n_clauses, (size_t)1));
    int max_var = f->max_var;
    std::vector<int32_t> dense;
    const size_t ext_before = f->ext_of.size();
    if (!f->ext_of.empty()) {                      // the frontier was renumbered when it was loaded: same map for its root
        dense.resize(3 * n_clauses + 1);
        if ((rc = map_in(f, clauses3, 3 * n_clauses, dense.data()))) return rc;
        clauses3 = dense.data();
    }
    if (!pack_host(clauses3, n_clauses, packed.data(), max_var)) {
        // (a root with ids beyond the codes under a frontier that needed no renumbering cannot be its ancestor)
        return fail(f->ext_of.empty() ? NDP_E_INVALID : NDP_E_UNSUPPORTED, "root literal magnitude exceeds 32766 (16-bit clause packing)");
    }
    const int old_max_var = f->max_var;
    f->root_c3.assign(clauses3, clauses3 + 3 * n_clauses);
    f->has_root = true;
    f->max_var = max_var;
    auto undo = [&](int code, const std::string& msg) {
        for (auto& r : f->root_indexes) {
            if (r.d_root) dev_free(r.d_root);
            if (r.d_hdr) dev_free(r.d_hdr);
            if (r.d_rows) dev_free(r.d_rows);
            if (r.d_occ_off) dev_free(r.d_occ_off);
            if (r.d_occ) dev_free(r.d_occ);
        }
        f->root_indexes.clear();
        if (f->sbuf[0]) { dev_free(f->sbuf[0]); f->sbuf[0] = nullptr; f->scap_slots[0] = 0; }
        f->root_c3.clear();
        f->has_root = false;
        f->states = false;
        f->max_var = old_max_var;
        while (f->ext_of.size() > ext_before && ext_before > 0) {   // variables only the refused root named
            f->dense_of.erase(f->ext_of.back());
            f->ext_of.pop_back();
        }
        return fail(code, msg);
    };
#define ATT_TRY(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return undo(NDP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
    const IndexRoot* ix = nullptr;
    if ((rc = get_root_index(f, f->device, &ix))) return undo(rc, g_err);
    const size_t slot_bytes = (size_t)ix->slot_words * 4;
    if (8 * slot_bytes > 227 * 1024) return undo(NDP_E_UNSUPPORTED, "root too large for the index engine's state slots");
    // 1. assignment table of every element from its choices, 2. state slots from the tables
    const size_t tw = (size_t)ix->table_words;
    std::vector<unsigned> tables(count * tw, 0u);
    for (size_t k = 0; k < count; ++k)
        for (size_t i = f->flat_off[k]; i < f->flat_off[k + 1]; ++i) {
            const int l = f->flat_choices[i];
            const unsigned v = (unsigned)std::abs(l);
            if (l == 0 || v > (unsigned)max_var) return undo(NDP_E_INVALID, "a choice names a variable the root does not have");
            tables[k * tw + (v >> 4)] |= (l > 0 ? 1u : 2u) << ((v & 15u) * 2u);
        }
    DevBuf d_tables, d_meta, d_scratch, d_n, d_expect, d_flag, d_ptr;
    cudaError_t ce;
    if ((ce = dev_alloc(&d_tables.p, std::max(tables.size(), (size_t)1) * 4)) != cudaSuccess ||
        (ce = dev_alloc(&d_meta.p, count * sizeof(BfsNode))) != cudaSuccess ||
        (ce = dev_alloc(&f->sbuf[0], count * slot_bytes)) != cudaSuccess)
        return undo(NDP_E_NOMEM, cudaGetErrorString(ce));
    f->scap_slots[0] = count;
    ATT_TRY(cudaMemcpy(d_tables.p, tables.data(), tables.size() * 4, cudaMemcpyHostToDevice));
    f->h2d_bytes += tables.size() * 4;
    ATT_TRY(cudaFuncSetAttribute(index_states_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * slot_bytes)));
    {
        const int wpb = 4;
        const int grid = (int)std::min<size_t>((count + wpb - 1) / wpb, (size_t)148 * 16);
        index_states_kernel<<<grid, wpb * 32, wpb * slot_bytes>>>(*ix, d_tables.as<unsigned>(), f->sbuf[0], (int)count,
                                                                  d_meta.as<BfsNode>());
        ATT_TRY(cudaGetLastError());
    }
    // 3. the proof: root filtered by each element's table must reproduce the element's clause list
    std::vector<int32_t> expect(count);
    for (size_t k = 0; k < count; ++k) expect[k] = f->entries[k].n;
    const size_t sstride = (size_t)pad_up((int)n_clauses + 1);
    const size_t chunk = std::max<size_t>(1, std::min(count, ((size_t)256 << 20) / (sstride * sizeof(pclause))));
    if ((ce = dev_alloc(&d_scratch.p, chunk * sstride * sizeof(pclause))) != cudaSuccess ||
        (ce = dev_alloc(&d_n.p, chunk * 4)) != cudaSuccess || (ce = dev_alloc(&d_expect.p, count * 4)) != cudaSuccess ||
        (ce = dev_alloc(&d_flag.p, 4)) != cudaSuccess || (ce = dev_alloc(&d_ptr.p, chunk * sizeof(unsigned*))) != cudaSuccess)
        return undo(NDP_E_NOMEM, cudaGetErrorString(ce));
    ATT_TRY(cudaMemcpy(d_expect.p, expect.data(), count * 4, cudaMemcpyHostToDevice));
    ATT_TRY(cudaMemset(d_flag.p, 0, 4));
    std::vector<const unsigned*> ptrs(chunk);
    for (size_t k0 = 0; k0 < count; k0 += chunk) {
        const size_t c = std::min(chunk, count - k0);
        for (size_t i = 0; i < c; ++i) ptrs[i] = f->sbuf[0] + (k0 + i) * (size_t)ix->slot_words;
        ATT_TRY(cudaMemcpy(d_ptr.p, ptrs.data(), c * sizeof(unsigned*), cudaMemcpyHostToDevice));
        const int grid = (int)std::min<size_t>((c + 3) / 4, (size_t)148 * 16);
        index_materialise_kernel<<<grid, 128>>>(*ix, d_ptr.as<const unsigned*>(), (int)c, d_scratch.as<pclause>(), sstride,
                                                d_n.as<int32_t>());
        clause_lists_equal_kernel<<<grid, 128>>>(d_scratch.as<pclause>(), sstride, d_n.as<int32_t>(),
                                                 f->buf[0] + k0 * f->stride, f->stride, d_expect.as<int32_t>() + k0, (int)c,
                                                 d_flag.as<unsigned>());
        ATT_TRY(cudaGetLastError());
    }
    unsigned mismatch = 0;
    ATT_TRY(cudaMemcpy(&mismatch, d_flag.p, 4, cudaMemcpyDeviceToHost));
    if (mismatch) return undo(NDP_E_INVALID, "the frontier is not this root filtered by its choices (different DIMACS file?)");
    std::vector<BfsNode> meta(count);
    ATT_TRY(cudaMemcpy(meta.data(), d_meta.p, count * sizeof(BfsNode), cudaMemcpyDeviceToHost));
    f->d2h_bytes += count * sizeof(BfsNode) + 4;
    for (size_t k = 0; k < count; ++k) {
        Entry& e = f->entries[k];
        e.buf = 0; e.slot = (uint32_t)k; e.kind = meta[k].kind; e.lit = meta[k].lit;   // e.n and the clause lists stay as loaded
    }
    f->states = true;
    f->materialised = true;   // buf[0] slot k already holds element k's clause list
    return NDP_OK;            // (scan replicas built before the call stay valid: the clause lists did not move)
#undef ATT_TRY
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
    int rc = ensure_materialised(f);   // a state frontier searched by the scan engine needs clause lists
    if (rc) return rc;
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
        CUDA_TRY(dev_alloc(&r.shard_buf, r.n_local * f->stride * sizeof(pclause)));
        for (size_t q = 0; q < r.n_local; ++q) {
            const size_t k = first + q * stride;
            CUDA_TRY(cudaMemcpyPeerAsync(r.shard_buf + q * f->stride, device, clause_slot(f, k), f->device,
                                         (size_t)f->entries[k].n * sizeof(pclause), 0));
        }
        CUDA_TRY(cudaStreamSynchronize(0));
    }
    for (size_t q = 0; q < r.n_local; ++q) {
        const size_t k = first + q * stride;
        h_base[q] = r.shard_buf ? r.shard_buf + q * f->stride : clause_slot(f, k);
        h_n[q] = f->entries[k].n;
        r.n_max = std::max(r.n_max, f->entries[k].n);
    }
    CUDA_TRY(dev_alloc((void**)&r.d_base, h_base.size() * sizeof(pclause*)));
    CUDA_TRY(dev_alloc(&r.d_n, h_n.size() * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy((void*)r.d_base, h_base.data(), h_base.size() * sizeof(pclause*), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_n, h_n.data(), h_n.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    f->replicas.push_back(r);
    *out = &f->replicas.back();
    return NDP_OK;
}

// ---- index DFS ------------------------------------------------------------------------------
bool engine_uses_index(const ndp_frontier* f) {
    if (!f->has_root) return false;   // explicit clause sets (resume): no common root to index
    return requested_engine() != NDP_ENGINE_SCAN;
}

struct IndexLaunchPlan {
    int warps_per_cta, ctas_per_sm, grid, smem_bytes;
    int groups;          // searches per warp (ndp_index_group.cuh): 1, 2 (16-lane groups) or 4 (8-lane groups)
    bool global_state;   // state slots in global memory (experiment knob NDP_DFS_GLOBAL_STATE, groups > 1 only)
    IndexParams p;
    int searchers() const { return grid * warps_per_cta * groups; }   // pieces in flight on the device
};

int plan_index(int device, const IndexRoot& ix, int max_var, IndexLaunchPlan& plan) {
    int sms = 0, smem_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    IndexParams& p = plan.p;
    p.ix = ix;
    p.trail_cap = std::max(max_var, 1);
    p.pend_words = (p.trail_cap + 31) / 32;
    size_t off = align16((size_t)ix.slot_words * 4);
    p.off_pend = (int)off;  off += align16((size_t)p.pend_words * 4);
    p.warp_bytes = (int)std::max(off, (size_t)16);
    if ((size_t)p.warp_bytes > (size_t)smem_optin)
        return fail(NDP_E_UNSUPPORTED, "clause count too large for the shared-memory bitmaps");
    const int sm_smem = 233472;
    // Searches per warp.  The kernel is bound by instruction issue and almost all of a node's instructions are
    // lane-independent, so several pieces per warp (lane groups walking the unit step in lockstep) multiply the
    // node rate -- as long as shared memory still holds enough pieces to keep a few warps per scheduler resident.
    // Groups need the plain root (B[2] summary, no {0,0,0} clause, no repeated variable in a clause).
    // Measured (profiles/r2_dfs_group_*): 32-bit CNF (2.1 KB slots, 104 pieces per SM) 1.53 -> 1.06 ms with 4 groups,
    // half the warp instructions per node; 48-bit (4.3 KB slots) no gain with 2 groups and a loss with 4 -- there the
    // kernel is bound by resident pieces / per-node latency, and the larger carve-out takes the index's L1 away.
    const int pieces_per_sm = (int)((size_t)sm_smem / ((size_t)p.warp_bytes + 256));
    int groups = ix.plain && pieces_per_sm >= 96 ? 4 : 1;
    // Larger slots: what bounds the kernel is pieces in flight / per-node latency, and shared memory holds too few
    // pieces (48-bit: 51 per SM, 64-bit: 29).  There the slots live in GLOBAL memory (L1/L2-served) under the lane-group
    // kernel: 48 warps x 4 groups = 192 pieces per SM.  48-bit exhaustive 600 -> 453 ms, 64-bit to the factors 50.4 -> 35.3 s
    // (32 warps) on one B200.
    bool global_state = ix.plain && pieces_per_sm < 96;
    if (const char* s = getenv("NDP_DFS_GLOBAL_STATE")) global_state = ix.plain && atoi(s) != 0;
    if (global_state) groups = 4;
    if (const char* s = getenv("NDP_DFS_GROUPS")) {
        const int g = atoi(s);
        if (ix.plain && (g == 1 || g == 2 || g == 4)) groups = g;
    }
    plan.global_state = groups > 1 && global_state;
    if (plan.global_state) {
        plan.groups = groups;
        plan.warps_per_cta = 4;
        plan.ctas_per_sm = 12;
        if (const char* s = getenv("NDP_INDEX_WARPS_PER_SM")) plan.ctas_per_sm = std::max(1, std::min(12, atoi(s) / 4));
        plan.smem_bytes = 0;
        plan.grid = sms * plan.ctas_per_sm;
        return NDP_OK;
    }
    int best_total = 0, best_w = 1, best_ctas = 1;
    int cap_warps = 48;   // resident warps per SM (above 32 the 40-register build of the kernel runs): 48-bit CNF 638 -> 600 ms
    if (const char* s = getenv("NDP_INDEX_WARPS_PER_SM")) cap_warps = std::max(1, std::min(48, atoi(s)));
    for (;;) {
        best_total = 0;
        for (int w : {4, 2, 1}) {
            size_t cta = (size_t)p.warp_bytes * w * groups;
            if (cta > (size_t)smem_optin) continue;
            int ctas = (int)std::min<size_t>(cap_warps / w, (size_t)sm_smem / (cta + 1024));
            if (ctas < 1) continue;
            if (w * ctas > best_total) { best_total = w * ctas; best_w = w; best_ctas = ctas; }
        }
        if (best_total > 0 || groups == 1) break;
        groups /= 2;      // the slots of `groups` searches do not fit one CTA: fewer searches per warp
    }
    if (best_total == 0) return fail(NDP_E_UNSUPPORTED, "clause count too large for the shared-memory bitmaps");
    plan.groups = groups;
    plan.warps_per_cta = best_w;
    plan.ctas_per_sm = best_ctas;
    plan.smem_bytes = p.warp_bytes * best_w * groups;
    plan.grid = sms * best_ctas;
    return NDP_OK;
}

int build_index_replica(ndp_frontier* f, int device, size_t first, size_t stride, const IndexRoot** ix_out,
                        ndp_frontier::IndexReplica** out) {
    int rc = get_root_index(f, device, ix_out);
    if (rc) return rc;
    const IndexRoot& ix = **ix_out;
    for (auto& r : f->index_replicas)
        if (r.device == device && r.first == first && r.stride == stride) { *out = &r; return NDP_OK; }
    ndp_frontier::IndexReplica r;
    r.device = device;
    r.first = first;
    r.stride = stride;
    const size_t total = f->entries.size();
    r.n_local = first < total ? (total - first + stride - 1) / stride : 0;
    const size_t nl1 = std::max(r.n_local, (size_t)1);
    for (size_t q = 0; q < r.n_local; ++q) r.mean_alive += (double)f->entries[first + q * stride].n;
    if (r.n_local) r.mean_alive /= (double)r.n_local;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(dev_alloc((void**)&r.d_slot_ptr, nl1 * sizeof(unsigned*)));
    CUDA_TRY(dev_alloc(&r.d_meta, nl1 * sizeof(BfsNode)));
    std::vector<const unsigned*> h_ptr(nl1, nullptr);
    std::vector<BfsNode> h_meta(nl1);
    if (f->states) {
        // the BFS already produced state slots: use them in place (home device) or copy the shard over
        if (device != f->device && r.n_local) {
            // gather the shard's slots into one buffer on the home device, then ONE peer copy
            const size_t bytes = r.n_local * (size_t)ix.slot_words * 4;
            CUDA_TRY(dev_alloc(&r.own_slots, bytes));
            std::vector<const unsigned*> src(r.n_local);
            for (size_t q = 0; q < r.n_local; ++q) {
                const Entry& e = f->entries[first + q * stride];
                src[q] = f->sbuf[e.buf] + (size_t)e.slot * ix.slot_words;
            }
            CUDA_TRY(cudaSetDevice(f->device));
            DevBuf d_src, d_stage;
            CUDA_TRY(dev_alloc(&d_src.p, r.n_local * sizeof(unsigned*)));
            CUDA_TRY(dev_alloc(&d_stage.p, bytes));
            CUDA_TRY(cudaMemcpy(d_src.p, src.data(), r.n_local * sizeof(unsigned*), cudaMemcpyHostToDevice));
            slot_gather_kernel<<<(int)std::min<size_t>((r.n_local + 3) / 4, (size_t)148 * 16), 128>>>(
                d_src.as<const unsigned*>(), (int)r.n_local, ix.slot_words, d_stage.as<unsigned>());
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyPeer(r.own_slots, device, d_stage.p, f->device, bytes));
            CUDA_TRY(cudaDeviceSynchronize());
            dev_free(d_src.p); d_src.p = nullptr;      // freed on the device that owns them
            dev_free(d_stage.p); d_stage.p = nullptr;
            CUDA_TRY(cudaSetDevice(device));
            f->h2d_bytes += r.n_local * sizeof(unsigned*);
        }
        for (size_t q = 0; q < r.n_local; ++q) {
            const Entry& e = f->entries[first + q * stride];
            h_ptr[q] = r.own_slots ? r.own_slots + q * ix.slot_words : f->sbuf[e.buf] + (size_t)e.slot * ix.slot_words;
            h_meta[q] = BfsNode{(uint32_t)q, e.n, e.kind, e.lit};
        }
        CUDA_TRY(cudaMemcpy(r.d_meta, h_meta.data(), nl1 * sizeof(BfsNode), cudaMemcpyHostToDevice));
    } else if (r.n_local) {
        // clause-slot frontier with a root: each subproblem's choices -> assignment table (walking the
        // tree on the device) -> state slot
        DevBuf d_tables, d_entry, d_tp, d_tl;
        const size_t tbytes = r.n_local * (size_t)ix.table_words * 4;
        CUDA_TRY(dev_alloc(&d_tables.p, std::max(tbytes, (size_t)4)));
        CUDA_TRY(cudaMemset(d_tables.p, 0, std::max(tbytes, (size_t)4)));
        std::vector<int32_t> h_entry_tree(total);
        for (size_t k = 0; k < total; ++k) h_entry_tree[k] = f->entries[k].tree;
        CUDA_TRY(dev_alloc(&d_entry.p, total * 4));
        CUDA_TRY(cudaMemcpy(d_entry.p, h_entry_tree.data(), total * 4, cudaMemcpyHostToDevice));
        f->h2d_bytes += total * 4;
        const int32_t *tp = f->d_tree_parent, *tl = f->d_tree_lit;
        if (device != f->device || !tp) {   // another GPU: ship the tree from the host copy
            if ((rc = ensure_host_tree(f))) return rc;
            CUDA_TRY(cudaSetDevice(device));
            const size_t tn = f->tree_parent.size();
            CUDA_TRY(dev_alloc(&d_tp.p, tn * 4));
            CUDA_TRY(dev_alloc(&d_tl.p, tn * 4));
            CUDA_TRY(cudaMemcpy(d_tp.p, f->tree_parent.data(), tn * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(d_tl.p, f->tree_lit.data(), tn * 4, cudaMemcpyHostToDevice));
            f->h2d_bytes += tn * 8;
            tp = d_tp.as<int32_t>();
            tl = d_tl.as<int32_t>();
        }
        const int threads = 128;
        const int blocks = (int)std::min<size_t>((r.n_local + threads - 1) / threads, 4096);
        prefix_table_kernel<<<blocks, threads>>>(d_entry.as<int32_t>(), tp, tl, first, stride, r.n_local,
                                                 ix.table_words, d_tables.as<unsigned>());
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(dev_alloc(&r.own_slots, r.n_local * (size_t)ix.slot_words * 4));
        const int wpb = 4;
        const int grid = (int)std::min<size_t>((r.n_local + wpb - 1) / wpb, (size_t)148 * 16);
        index_states_kernel<<<grid, wpb * 32, wpb * ix.slot_words * 4>>>(ix, d_tables.as<unsigned>(), r.own_slots,
                                                                        (int)r.n_local, r.d_meta);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        for (size_t q = 0; q < r.n_local; ++q) h_ptr[q] = r.own_slots + q * ix.slot_words;
    }
    CUDA_TRY(cudaMemcpy((void*)r.d_slot_ptr, h_ptr.data(), nl1 * sizeof(unsigned*), cudaMemcpyHostToDevice));
    f->h2d_bytes += nl1 * (sizeof(unsigned*) + sizeof(BfsNode));
    f->index_replicas.push_back(r);
    *out = &f->index_replicas.back();
    return NDP_OK;
}

// ---- DFS-order split of a narrow shard (ndp_split.cuh) ----------------------------------------
struct SplitRun {
    bool active = false;
    DevBuf ptr[2], meta[2], orig[2], tnode[2], pre_n[2], pre_e[2], sbuf[2];
    DevBuf children, parents, t_parent, t_lit, t_off, t_len, log, tail_n, tail_e, state;
    SplitParams sp;
    SplitState st;
    size_t n_pieces = 0;
    int side = 0;       // which SplitLevel holds the final pieces
};

// How many pieces the shard should be cut into (0 = leave the subproblems whole).
// Automatic policy.  WHEN: the shard has fewer units than pieces-per-warp x resident warps AND the cut can
// pay for itself (a split level costs ~0.1-1 ms): either the shard cannot even occupy a quarter of the warps,
// or its subproblems are still big -- more than half of the root's clauses alive on average, which for these
// CNFs means thousands of nodes each.  (32-bit -q 16384 shared out over 8 ranks is 2048 subproblems of ~380
// nodes: searched whole in 0.87 ms, 0.99 ms with a cut; -q 1024, ~7800 nodes each: 3.8 ms cut, 17.9 ms whole.)
// HOW FINE: two pieces per warp keep the device busy; searches that run for seconds have heavy-tailed piece
// sizes, so large CNFs (deep subtrees) are cut finer -- the tail is what bounds an early-exit run.
// n_root_clauses = 0: the coarse cut that shares a narrow frontier out over several devices (each device then
// refines its share).  mean_alive = average live clauses per unit of the shard.
size_t split_target(size_t n_sub, int n_warps, int n_root_clauses, double mean_alive, int groups = 1) {
    long long mode = g_split;
    if (mode < 0)
        if (const char* e = getenv("NDP_SPLIT")) mode = !strcmp(e, "off") ? 0 : atoll(e);
    if (mode == 0 || n_sub == 0) return 0;
    if (mode > 0) return (size_t)mode > n_sub ? (size_t)mode : 0;
    // pieces per searcher (n_warps counts searchers: warps x lane groups).  Measured on the 48-bit CNF, 28416 searchers
    // per B200: 2 / 4 / 8 / 16 pieces each -> cut 4 / 10 / 17 / 34 ms, search 502 / 453 / 435 / 406 ms on one GPU; with the
    // work shared by 8 GPUs the cut costs the same and the tail is 8 x shorter, so 4 is the better compromise.
    size_t per_warp = n_root_clauses > 16384 ? 32 : n_root_clauses > 8192 ? 4 : groups > 1 ? 1 : 2;
    if (n_root_clauses > 0)
        if (const char* e = getenv("NDP_SPLIT_PER_WARP")) per_warp = (size_t)std::max(1, atoi(e));
    if (n_sub >= per_warp * (size_t)n_warps) return 0;
    const bool starved = 4 * n_sub < (size_t)n_warps / (size_t)std::max(groups, 1);   // fewer units than a quarter of the WARPS
    const bool deep = n_root_clauses <= 0 || mean_alive > 0.5 * (double)n_root_clauses;
    return starved || deep ? per_warp * (size_t)n_warps : 0;
}

int run_split_once(int device, const IndexRoot& ix, const ndp_frontier::IndexReplica& rep, size_t target, SplitRun& run,
                   ndp_stats* stats) {
    const size_t n_sub = rep.n_local;
    int sms = 0, coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    SplitParams& sp = run.sp;
    sp.ix = ix;
    sp.chain_cap = 128;
    sp.max_levels = 64;
    if (const char* e = getenv("NDP_SPLIT_CHAIN")) sp.chain_cap = std::max(1, std::min(4096, atoi(e)));
    if (const char* e = getenv("NDP_SPLIT_LEVELS")) sp.max_levels = std::max(1, atoi(e));
    sp.target = (unsigned)std::min<size_t>(target, 1u << 22);
    sp.off_chain = (int)align16((size_t)ix.slot_words * 4);
    sp.warp_bytes = sp.off_chain + (int)align16((size_t)sp.chain_cap * 2);
    const int warps = (int)std::min<size_t>(32, ((size_t)200 * 1024) / (size_t)sp.warp_bytes);
    if (!coop || warps < 1) return NDP_OK;   // cannot co-schedule the grid: search the subproblems whole
    const void* skern = ix.plain ? (const void*)dfs_split_kernel<true> : (const void*)dfs_split_kernel<false>;
    CUDA_TRY(cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * sp.warp_bytes));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, skern, warps * 32,
                                                           (size_t)warps * sp.warp_bytes));
    if (per_sm < 1) return NDP_OK;
    const size_t cap = 2 * (size_t)sp.target;                 // pieces per level (children of a level: <= cap)
    const size_t cap_tree = n_sub + 24 * cap;
    const size_t cap_log = std::max<size_t>((size_t)32 << 20, cap * (size_t)sp.chain_cap);   // shorts; a level needs m * chain_cap free
    const size_t slot_bytes = (size_t)ix.slot_words * 4;
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(dev_alloc(&run.ptr[b].p, cap * sizeof(unsigned*)));
        CUDA_TRY(dev_alloc(&run.meta[b].p, cap * sizeof(BfsNode)));
        CUDA_TRY(dev_alloc(&run.orig[b].p, cap * 4));
        CUDA_TRY(dev_alloc(&run.tnode[b].p, cap * 4));
        CUDA_TRY(dev_alloc(&run.pre_n[b].p, cap * 8));
        CUDA_TRY(dev_alloc(&run.pre_e[b].p, cap * 8));
        CUDA_TRY(dev_alloc(&run.sbuf[b].p, cap * slot_bytes));
        sp.lv[b].ptr = run.ptr[b].as<const unsigned*>();
        sp.lv[b].meta = run.meta[b].as<BfsNode>();
        sp.lv[b].orig = run.orig[b].as<int32_t>();
        sp.lv[b].tnode = run.tnode[b].as<int32_t>();
        sp.lv[b].pre_nodes = run.pre_n[b].as<unsigned long long>();
        sp.lv[b].pre_evals = run.pre_e[b].as<unsigned long long>();
        sp.sbuf[b] = run.sbuf[b].as<unsigned>();
    }
    CUDA_TRY(dev_alloc(&run.children.p, cap * sizeof(BfsChild)));
    CUDA_TRY(dev_alloc(&run.parents.p, (cap / 2 + 1) * sizeof(SplitParent)));
    CUDA_TRY(dev_alloc(&run.t_parent.p, cap_tree * 4));
    CUDA_TRY(dev_alloc(&run.t_lit.p, cap_tree * 4));
    CUDA_TRY(dev_alloc(&run.t_off.p, cap_tree * 4));
    CUDA_TRY(dev_alloc(&run.t_len.p, cap_tree * 4));
    CUDA_TRY(dev_alloc(&run.log.p, cap_log * sizeof(short)));
    CUDA_TRY(dev_alloc(&run.tail_n.p, n_sub * 8));
    CUDA_TRY(dev_alloc(&run.tail_e.p, n_sub * 8));
    CUDA_TRY(dev_alloc(&run.state.p, sizeof(SplitState)));
    sp.children = run.children.as<BfsChild>();
    sp.parents = run.parents.as<SplitParent>();
    sp.t_parent = run.t_parent.as<int32_t>();
    sp.t_lit = run.t_lit.as<int32_t>();
    sp.t_seg_off = run.t_off.as<unsigned>();
    sp.t_seg_len = run.t_len.as<unsigned>();
    sp.log = run.log.as<short>();
    sp.tail_nodes = run.tail_n.as<unsigned long long>();
    sp.tail_evals = run.tail_e.as<unsigned long long>();
    sp.st = run.state.as<SplitState>();
    SplitState& st = run.st;
    memset(&st, 0, sizeof st);
    st.m = (int)n_sub;
    st.tree_count = n_sub;
    st.cap_pieces = cap;
    st.cap_tree = cap_tree;
    st.cap_log = cap_log;
    CUDA_TRY(cudaMemcpyAsync(sp.st, &st, sizeof st, cudaMemcpyHostToDevice, 0));
    split_init_kernel<<<(int)std::min<size_t>((n_sub + 255) / 256, 1024), 256>>>(sp, rep.d_slot_ptr, rep.d_meta, (int)n_sub);
    CUDA_TRY(cudaGetLastError());
    void* args[] = {&sp};
    CUDA_TRY(cudaLaunchCooperativeKernel(skern, dim3(sms), dim3(warps * 32), args,
                                         (size_t)warps * sp.warp_bytes, 0));
    CUDA_TRY(cudaMemcpy(&st, sp.st, sizeof st, cudaMemcpyDeviceToHost));
    run.active = true;
    run.n_pieces = (size_t)st.m;
    run.side = st.cur;
    if (stats) { stats->launches += 2; stats->h2d_bytes += sizeof st; stats->d2h_bytes += sizeof st; }
    if (tracing())
        fprintf(stderr, "[ndp_split] %zu subproblems -> %zu pieces in %d levels (exit %d), tree %llu, log %llu\n", n_sub,
                run.n_pieces, st.levels_done, st.exit_reason, st.tree_count, st.log_count);
    return NDP_OK;
}

// The cut with a target that fits the device: the piece arenas are 4 x target state slots, and a large CNF on a
// busy device may not have room for the target the policy asks for.  Out of memory halves the target (fewer, larger
// pieces: same results, a longer tail) instead of failing the search; below the shard's own size there is no cut.
int run_split(int device, const IndexRoot& ix, const ndp_frontier::IndexReplica& rep, size_t target, SplitRun& run,
              ndp_stats* stats) {
    for (;;) {
        if (target <= rep.n_local) return NDP_OK;         // no room for any cut: the units are searched whole
        int rc = run_split_once(device, ix, rep, target, run, stats);
        if (rc != NDP_E_NOMEM) return rc;
        cudaGetLastError();                                // clear the sticky allocation error
        run = SplitRun();                                  // releases what the failed attempt had allocated
        target /= 2;
        if (tracing()) fprintf(stderr, "[ndp_split] out of device memory: retrying with a target of %zu pieces\n", target);
    }
}

// One shard on one device.  `shared_best`/`peer_best`: early-exit words (this device's + peers').
int dfs_on_device(ndp_frontier* f, int device, size_t first, size_t stride, int early_exit,
                  unsigned long long* shared_best, unsigned long long* const* peer_best, int n_peers,
                  unsigned long long best_init, uint8_t* verdict, size_t* winner, std::vector<int32_t>* model,
                  uint64_t* nodes, uint64_t* evals, ndp_stats* stats) {
    const bool use_index = engine_uses_index(f);
    int rc;
    const double t_begin = now_ms();
    double t_replica = 0, t_launched = 0, t_kernel_done = 0, t_results = 0;
    ndp_frontier::Replica* rep = nullptr;
    ndp_frontier::IndexReplica* irep = nullptr;
    const IndexRoot* ix = nullptr;
    if (use_index) rc = build_index_replica(f, device, first, stride, &ix, &irep);
    else rc = build_replica(f, device, first, stride, &rep);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    t_replica = now_ms();
    const size_t n_sub = use_index ? irep->n_local : rep->n_local;   // subproblems of this shard
    if (winner) *winner = SIZE_MAX;
    if (n_sub == 0) return NDP_OK;
    if (n_sub > 0x7ffffff0ull) return fail(NDP_E_UNSUPPORTED, "shard larger than the 32-bit work counter");

    DfsLaunchPlan plan;
    IndexLaunchPlan iplan;
    int n_warps, model_stride;
    if (use_index) {
        if ((rc = plan_index(device, *ix, f->max_var, iplan))) return rc;
        n_warps = iplan.searchers();    // searches in flight: every per-"warp" array below is per searcher
        model_stride = iplan.p.trail_cap;
    } else {
        if ((rc = plan_dfs(device, rep->n_max, f->max_var, plan))) return rc;
        n_warps = plan.grid * plan.warps_per_cta;
        plan.p.model_stride = plan.p.trail_cap;
        model_stride = plan.p.model_stride;
    }

    cudaEvent_t ev0, ev1;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    struct EvFree { cudaEvent_t a, b; ~EvFree() { cudaEventDestroy(a); cudaEventDestroy(b); } } ev_free{ev0, ev1};
    CUDA_TRY(cudaEventRecord(ev0, 0));

    // index engine: a shard too narrow for the device is first cut into pieces in DFS order
    SplitRun split;
    if (use_index) {
        const size_t target = split_target(n_sub, n_warps, ix->n_root, irep->mean_alive, iplan.groups);
        if (target > n_sub && (rc = run_split(device, *ix, *irep, target, split, stats))) return rc;
    }
    const size_t nl = split.active ? split.n_pieces : n_sub;         // units the DFS kernel pulls

    DevBuf b_verdict, b_nodes, b_evals, b_reb, b_next, b_best, b_bestp, b_msub, b_mlen, b_models, b_scratch, b_snaps;
    int snap_depth = 4;
    if (const char* e = getenv("NDP_DFS_SNAPS")) snap_depth = std::max(0, std::min(32, atoi(e)));
    if (use_index && iplan.groups > 1) snap_depth = std::min(snap_depth, 32 / iplan.groups);   // lane k of a group holds |A| of snapshot k
    const size_t nl1 = std::max(nl, (size_t)1);
    CUDA_TRY(dev_alloc(&b_verdict.p, nl1));
    CUDA_TRY(dev_alloc(&b_nodes.p, nl1 * 8));
    CUDA_TRY(dev_alloc(&b_evals.p, nl1 * 8));
    CUDA_TRY(dev_alloc(&b_reb.p, nl1 * 4));
    CUDA_TRY(dev_alloc(&b_next.p, 4));
    CUDA_TRY(dev_alloc(&b_msub.p, (size_t)n_warps * 8));
    CUDA_TRY(dev_alloc(&b_mlen.p, (size_t)n_warps * 4));
    CUDA_TRY(dev_alloc(&b_models.p, (size_t)n_warps * model_stride * 4));
    if (use_index) {
        CUDA_TRY(dev_alloc(&b_scratch.p, (size_t)n_warps * model_stride * sizeof(short)));   // trails
        if (snap_depth > 0 &&
            dev_alloc(&b_snaps.p, (size_t)n_warps * snap_depth * ix->slot_words * 4) != cudaSuccess) {
            cudaGetLastError();   // no room for snapshots: pending siblings are re-derived from the trail
            b_snaps.p = nullptr;
            snap_depth = 0;
        }
        CUDA_TRY(dev_alloc(&b_bestp.p, n_sub * 4));
        CUDA_TRY(cudaMemsetAsync(b_bestp.p, 0xff, n_sub * 4, 0));
    } else if (!plan.smem_w) {
        CUDA_TRY(dev_alloc(&b_scratch.p, (size_t)n_warps * plan.p.w_cap * sizeof(pclause)));
    }
    unsigned long long* d_best = shared_best;
    if (!d_best) {
        CUDA_TRY(dev_alloc(&b_best.p, 8));
        d_best = b_best.as<unsigned long long>();
        CUDA_TRY(cudaMemcpyAsync(d_best, &best_init, 8, cudaMemcpyHostToDevice, 0));
    }
    CUDA_TRY(cudaMemsetAsync(b_verdict.p, NDP_NOT_RUN, nl1, 0));
    CUDA_TRY(cudaMemsetAsync(b_next.p, 0, 4, 0));
    CUDA_TRY(cudaMemsetAsync(b_msub.p, 0xff, (size_t)n_warps * 8, 0));

    if (nl == 0) {
        // every piece died inside the split: nothing left to search
    } else if (use_index) {
        IndexParams& p = iplan.p;
        if (split.active) {
            const SplitLevel& L = split.sp.lv[split.side];
            p.slot_ptr = L.ptr;
            p.meta = L.meta;
            p.orig = L.orig;
        } else {
            p.slot_ptr = irep->d_slot_ptr;
            p.meta = irep->d_meta;
            p.orig = nullptr;
        }
        p.best_piece = b_bestp.as<unsigned int>();
        p.first = first; p.stride = stride; p.n_local = nl;
        p.max_steps = 0;
        if (const char* e = getenv("NDP_PROFILE_MAX_STEPS")) p.max_steps = (unsigned)std::max(0ll, atoll(e));   // profiling knob (group kernel)
        if (const char* e = getenv("NDP_PROFILE_MAX_UNITS"))   // profiling knob: search only the first units (the rest stay NOT_RUN)
            p.n_local = std::min<unsigned long long>(nl, (unsigned long long)std::max(1ll, atoll(e)));
        p.next = b_next.as<unsigned int>();
        progress_arm(device, nl, &p.progress);
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
        p.snaps = b_snaps.as<unsigned>();
        p.snap_depth = snap_depth;
        p.models = b_models.as<int32_t>();
        // the common case (B[2] summary, no {0,0,0} clause, no variable twice in a clause) runs the kernel
        // with those tests compiled out
        const bool dense = iplan.warps_per_cta * iplan.ctas_per_sm > 32;   // more than 32 warps per SM: the 40-register build
        DevBuf b_gstate;
        p.gstate = nullptr;
        if (iplan.global_state) {
            CUDA_TRY(dev_alloc(&b_gstate.p, (size_t)n_warps * p.warp_bytes));
            p.gstate = b_gstate.as<unsigned char>();
        }
        auto kern = iplan.global_state ? (iplan.groups == 4 ? (dense ? dfs_index_group_kernel<8, true, 12> : dfs_index_group_kernel<8, true>)
                                                            : (dense ? dfs_index_group_kernel<16, true, 12> : dfs_index_group_kernel<16, true>))
                    : iplan.groups == 4 ? dfs_index_group_kernel<8> : iplan.groups == 2 ? dfs_index_group_kernel<16>
                    : ix->plain ? (dense ? dfs_index_kernel<true, 12> : dfs_index_kernel<true, 8>)
                                : (dense ? dfs_index_kernel<false, 12> : dfs_index_kernel<false, 8>);
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, iplan.smem_bytes));
        {
            // leave L1 everything the resident CTAs do not need: the root index is read through it
            const size_t need = (size_t)iplan.ctas_per_sm * (iplan.smem_bytes + 1024);
            int pct = (int)std::min<size_t>(100, (need * 100 + 233471) / 233472);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        kern<<<iplan.grid, iplan.warps_per_cta * 32, iplan.smem_bytes>>>(p);
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
        kern<<<plan.grid, plan.warps_per_cta * 32, plan.smem_bytes>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev1, 0));
    t_launched = now_ms();
    {
        cudaError_t se = cudaEventSynchronize(ev1);
        progress_disarm(device);
        CUDA_TRY(se);
    }
    t_kernel_done = now_ms();
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);

    // ---- per-piece results -> per-subproblem results
    std::vector<uint8_t> h_v(nl1);
    std::vector<unsigned long long> h_nodes(nl1), h_evals(nl1);
    std::vector<unsigned> h_reb(nl1);
    size_t d2h = 0;
    if (nl) {
        CUDA_TRY(cudaMemcpy(h_v.data(), b_verdict.p, nl, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_nodes.data(), b_nodes.p, nl * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_evals.data(), b_evals.p, nl * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_reb.data(), b_reb.p, nl * 4, cudaMemcpyDeviceToHost));
        d2h += nl * (1 + 8 + 8 + 4);
    }
    if (const char* dump = getenv("NDP_DUMP_PIECES")) {   // diagnostic: live clauses at the start of each unit vs nodes it took
        std::vector<BfsNode> h_meta(nl1);
        const BfsNode* d_meta = use_index ? (split.active ? split.sp.lv[split.side].meta : irep->d_meta) : nullptr;
        if (d_meta && nl && cudaMemcpy(h_meta.data(), d_meta, nl * sizeof(BfsNode), cudaMemcpyDeviceToHost) == cudaSuccess)
            if (FILE* fp = fopen(dump, "w")) {
                for (size_t q = 0; q < nl; ++q) fprintf(fp, "%d,%d,%llu\n", h_meta[q].n, h_meta[q].kind, h_nodes[q]);
                fclose(fp);
            }
    }
    std::vector<uint8_t> sub_v;
    std::vector<unsigned long long> sub_nodes, sub_evals;
    std::vector<unsigned> h_bestp;
    uint64_t tr = 0;
    for (size_t q = 0; q < nl; ++q) tr += h_reb[q];
    if (split.active) {
        std::vector<int32_t> h_orig(nl1);
        std::vector<unsigned long long> h_pre_n(nl1), h_pre_e(nl1), h_tail_n(n_sub), h_tail_e(n_sub);
        h_bestp.resize(n_sub);
        const SplitLevel& L = split.sp.lv[split.side];
        if (nl) {
            CUDA_TRY(cudaMemcpy(h_orig.data(), L.orig, nl * 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(h_pre_n.data(), L.pre_nodes, nl * 8, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(h_pre_e.data(), L.pre_evals, nl * 8, cudaMemcpyDeviceToHost));
        }
        CUDA_TRY(cudaMemcpy(h_tail_n.data(), split.sp.tail_nodes, n_sub * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_tail_e.data(), split.sp.tail_evals, n_sub * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_bestp.data(), b_bestp.p, n_sub * 4, cudaMemcpyDeviceToHost));
        d2h += nl * 20 + n_sub * 20;
        sub_v.assign(n_sub, NDP_UNSAT);
        sub_nodes.assign(n_sub, 0);
        sub_evals.assign(n_sub, 0);
        // nodes(subproblem) = pieces up to and including the first one with a model, each with the
        // interior nodes charged to it; an UNSAT subproblem adds the dead ends after its last piece
        for (size_t pc = 0; pc < nl; ++pc) {
            const size_t o = (size_t)h_orig[pc];
            const unsigned w = h_bestp[o];
            if (w != 0xffffffffu && pc > (size_t)w) continue;
            sub_nodes[o] += h_pre_n[pc] + h_nodes[pc];
            sub_evals[o] += h_pre_e[pc] + h_evals[pc];
            if (h_v[pc] == NDP_NOT_RUN && sub_v[o] == NDP_UNSAT) sub_v[o] = NDP_NOT_RUN;
        }
        for (size_t o = 0; o < n_sub; ++o) {
            if (h_bestp[o] != 0xffffffffu) sub_v[o] = NDP_SAT;
            else if (sub_v[o] == NDP_UNSAT) { sub_nodes[o] += h_tail_n[o]; sub_evals[o] += h_tail_e[o]; }
        }
    } else {
        sub_v.swap(h_v);
        sub_nodes.swap(h_nodes);
        sub_evals.swap(h_evals);
    }
    size_t win = SIZE_MAX, win_q = SIZE_MAX;
    uint64_t tn = 0, te = 0;
    for (size_t q = 0; q < n_sub; ++q) {
        const size_t g = first + q * stride;
        if (verdict) verdict[g] = sub_v[q];
        if (nodes) nodes[g] = sub_nodes[q];
        if (evals) evals[g] = sub_evals[q];
        tn += sub_nodes[q]; te += sub_evals[q];
        if (sub_v[q] == NDP_SAT && win == SIZE_MAX) { win = g; win_q = q; }
    }
    if (winner) *winner = win;
    t_results = now_ms();
    if (win != SIZE_MAX && model) {
        // the piece (= the subproblem itself without a split) whose model is the reference's
        const unsigned long long win_piece = split.active ? (unsigned long long)h_bestp[win_q]
                                             : (use_index ? (unsigned long long)win_q : (unsigned long long)win);
        std::vector<unsigned long long> h_msub(n_warps);
        std::vector<int32_t> h_mlen(n_warps);
        CUDA_TRY(cudaMemcpy(h_msub.data(), b_msub.p, (size_t)n_warps * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_mlen.data(), b_mlen.p, (size_t)n_warps * 4, cudaMemcpyDeviceToHost));
        d2h += (size_t)n_warps * 12;
        model->clear();
        bool found = false;
        for (int w = 0; w < n_warps && !found; ++w)
            if (h_msub[w] == win_piece) {
                found = true;
                if ((rc = frontier_choices(f, win, *model))) return rc;   // choices_k ++ dfs choices (NDP:1430-1431)
                CUDA_TRY(cudaSetDevice(device));                          // (the tree lives on the frontier's home device)
                if (split.active) {
                    // ... ++ the choices the split made on the way to the winning piece
                    const int cap = std::max(f->max_var, 1) + 8;
                    int32_t leaf = -1, start = 0;
                    CUDA_TRY(cudaMemcpy(&leaf, split.sp.lv[split.side].tnode + win_piece, 4, cudaMemcpyDeviceToHost));
                    DevBuf d_out, d_start;
                    CUDA_TRY(dev_alloc(&d_out.p, (size_t)cap * 4));
                    CUDA_TRY(dev_alloc(&d_start.p, 4));
                    split_path_kernel<<<1, 1>>>(split.sp.t_parent, split.sp.t_lit, split.sp.t_seg_off, split.sp.t_seg_len,
                                                split.sp.log, leaf, d_out.as<int32_t>(), cap, d_start.as<int32_t>());
                    CUDA_TRY(cudaGetLastError());
                    CUDA_TRY(cudaMemcpy(&start, d_start.p, 4, cudaMemcpyDeviceToHost));
                    if (start < 0) return fail(NDP_E_CUDA, "split path longer than the variable count");
                    std::vector<int32_t> path((size_t)(cap - start));
                    if (!path.empty())
                        CUDA_TRY(cudaMemcpy(path.data(), d_out.as<int32_t>() + start, path.size() * 4, cudaMemcpyDeviceToHost));
                    model->insert(model->end(), path.begin(), path.end());
                    d2h += 8 + path.size() * 4;
                    if (stats) stats->launches += 1;
                }
                std::vector<int32_t> tail((size_t)h_mlen[w]);
                if (!tail.empty())
                    CUDA_TRY(cudaMemcpy(tail.data(), b_models.as<int32_t>() + (size_t)w * model_stride,
                                        tail.size() * 4, cudaMemcpyDeviceToHost));
                model->insert(model->end(), tail.begin(), tail.end());
                d2h += tail.size() * 4;
            }
        // not found: a lower SAT index was already known (exit_flag hint / a peer device), so this
        // shard's own winner never parked its trail -- the job's model comes from that other shard
        if (!found) model->clear();
    }
    if (tracing())
        fprintf(stderr, "[ndp_dfs] dev %d: replica %.2f ms | setup+launch %.2f ms | wait %.2f ms (events %.2f ms) | results %.2f ms | "
                        "model %.2f ms\n", device, t_replica - t_begin, t_launched - t_replica, t_kernel_done - t_launched, (double)ms,
                t_results - t_kernel_done, now_ms() - t_results);
    if (stats) {
        stats->h2d_bytes += 8 + 4;                                   // early-exit word + work counter
        stats->d2h_bytes += d2h;
        stats->nodes += tn;
        stats->clause_evals += te;
        stats->rebuilds += tr;
        stats->kernel_ms = std::max(stats->kernel_ms, (double)ms);
        stats->launches += nl ? 1 : 0;
    }
    return NDP_OK;
}

// The frontier on n_gpus devices of this process: static interleaved split k -> k mod n_gpus, one host
// thread per device, peer-visible early-exit words.  nodes / evals (may be null): per-element counters.
int batch_core(ndp_frontier* f, int n_gpus, int early_exit, uint8_t* verdict, size_t* winner, int32_t* model,
               size_t model_cap, size_t* model_len, uint64_t* nodes, uint64_t* evals, ndp_stats* stats) {
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
    // peer-visible early-exit words: one per device; a kernel atomically lowers the word of every device it
    // can actually reach (peer access enabled), the others learn the winner from the host after the join
    struct BestWords {
        std::vector<unsigned long long*> p;
        std::vector<int> dev;
        ~BestWords() { for (size_t i = 0; i < p.size(); ++i) if (p[i]) { cudaSetDevice(dev[i]); cudaFree(p[i]); } }
    } words;
    words.p.assign(n_gpus, nullptr);
    words.dev = devs;
    std::vector<unsigned long long*>& best = words.p;
    std::vector<std::vector<unsigned long long*>> reach(n_gpus);   // reach[i]: words device i may write (its own first)
    std::vector<std::vector<char>> peer_ok(n_gpus, std::vector<char>(n_gpus, 0));
    const unsigned long long none = ~0ull;
    for (int i = 0; i < n_gpus; ++i) {
        CUDA_TRY(cudaSetDevice(devs[i]));
        for (int j = 0; j < n_gpus; ++j)
            if (i != j) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess) { cudaGetLastError(); can = 0; }
                if (can) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                    if (e != cudaSuccess) { cudaGetLastError(); can = 0; }
                }
                peer_ok[i][j] = (char)can;
            }
        CUDA_TRY(cudaMalloc(&best[i], 8));
        CUDA_TRY(cudaMemcpy(best[i], &none, 8, cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < n_gpus; ++i) {
        reach[i].push_back(best[i]);
        for (int j = 0; j < n_gpus; ++j)
            if (i != j && peer_ok[i][j]) reach[i].push_back(best[j]);
    }
    // build replicas serially (peer copies from the home device), then search concurrently
    const bool use_index = engine_uses_index(f);
    for (int i = 0; i < n_gpus; ++i) {
        if (use_index) {
            const IndexRoot* ix = nullptr;
            ndp_frontier::IndexReplica* rep = nullptr;
            if ((rc = build_index_replica(f, devs[i], (size_t)i, (size_t)n_gpus, &ix, &rep))) return rc;
        } else {
            ndp_frontier::Replica* rep = nullptr;
            if ((rc = build_replica(f, devs[i], (size_t)i, (size_t)n_gpus, &rep))) return rc;
        }
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
            rcs[i] = dfs_on_device(f, devs[i], (size_t)i, (size_t)n_gpus, early_exit, best[i], reach[i].data(),
                                   (int)reach[i].size(), none, verdict, &wins[i], &models[i], nodes, evals, &st[i]);
            if (rcs[i]) errs[i] = g_err;
        });
    for (auto& t : th) t.join();
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


// Pieces of a split as a frontier of their own: element p = piece p (state slots in the split's arena,
// no choices), so that batch_core / dfs_on_device can shard them over devices or ranks like subproblems.
// Per-piece results are then folded back per subproblem exactly as in dfs_on_device.
struct PieceSet {
    std::unique_ptr<ndp_frontier> pf;                    // borrows the split's arena and the parent's root indexes
    ndp_frontier* parent = nullptr;
    std::vector<int32_t> orig;                           // [pieces] subproblem a piece belongs to
    std::vector<unsigned long long> pre_n, pre_e;        // [pieces] split-interior nodes / evals charged to the piece
    std::vector<unsigned long long> tail_n, tail_e;      // [subproblems] dead ends after the last piece
    size_t np = 0;
    // hand the borrowed parts back (the arena belongs to the split; root indexes built for new devices move to the parent)
    void release() {
        if (!pf) return;
        pf->sbuf[0] = nullptr;
        for (size_t i = pf->root_borrowed; i < pf->root_indexes.size(); ++i) parent->root_indexes.push_back(pf->root_indexes[i]);
        pf->root_borrowed = pf->root_indexes.size();
        pf.reset();
    }
    ~PieceSet() { release(); }
};

int make_piece_set(ndp_frontier* f, const IndexRoot& ix, SplitRun& split, PieceSet& ps) {
    const size_t n_sub = f->entries.size(), np = split.n_pieces;
    const SplitLevel& L = split.sp.lv[split.side];
    ps.parent = f;
    ps.np = np;
    const size_t np1 = std::max(np, (size_t)1);
    std::vector<const unsigned*> h_ptr(np1);
    std::vector<BfsNode> h_meta(np1);
    ps.orig.assign(np1, 0);
    ps.pre_n.assign(np1, 0); ps.pre_e.assign(np1, 0);
    ps.tail_n.assign(n_sub, 0); ps.tail_e.assign(n_sub, 0);
    CUDA_TRY(cudaSetDevice(f->device));
    if (np) {
        CUDA_TRY(cudaMemcpy(h_ptr.data(), L.ptr, np * sizeof(unsigned*), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h_meta.data(), L.meta, np * sizeof(BfsNode), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(ps.orig.data(), L.orig, np * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(ps.pre_n.data(), L.pre_nodes, np * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(ps.pre_e.data(), L.pre_evals, np * 8, cudaMemcpyDeviceToHost));
    }
    CUDA_TRY(cudaMemcpy(ps.tail_n.data(), split.sp.tail_nodes, n_sub * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(ps.tail_e.data(), split.sp.tail_evals, n_sub * 8, cudaMemcpyDeviceToHost));
    // every surviving piece of the last level sits in the arena that level wrote (LEAF pieces have no slot)
    unsigned* arena = split.sp.sbuf[split.side];
    ps.pf.reset(new ndp_frontier);
    ndp_frontier* pf = ps.pf.get();
    pf->device = f->device;
    pf->max_var = f->max_var;
    pf->states = true;
    pf->has_root = true;
    pf->root_c3 = f->root_c3;
    pf->root_indexes = f->root_indexes;      // same root: share the per-device indexes already built
    pf->root_borrowed = pf->root_indexes.size();
    pf->host_tree_ready = true;
    pf->sbuf[0] = arena;
    pf->scap_slots[0] = 0;
    pf->entries.reserve(np);
    for (size_t pc = 0; pc < np; ++pc) {
        uint32_t slot = 0;
        if (h_ptr[pc]) {
            const size_t off = (size_t)(h_ptr[pc] - arena);
            if (h_ptr[pc] < arena || off % (size_t)ix.slot_words) return fail(NDP_E_CUDA, "piece slot outside the split arena");
            slot = (uint32_t)(off / (size_t)ix.slot_words);
        }
        pf->entries.push_back(Entry{0, slot, h_meta[pc].n, -1, h_meta[pc].kind, h_meta[pc].lit});
    }
    return NDP_OK;
}

// choices_k ++ the split's path to piece `pwin` of subproblem `win` (the caller appends the DFS tail)
int model_prefix(ndp_frontier* f, SplitRun& split, size_t win, size_t pwin, std::vector<int32_t>& m, ndp_stats* stats) {
    int rc = frontier_choices(f, win, m);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(f->device));
    const SplitLevel& L = split.sp.lv[split.side];
    const int cap = std::max(f->max_var, 1) + 8;
    int32_t leaf = -1, start = 0;
    CUDA_TRY(cudaMemcpy(&leaf, L.tnode + pwin, 4, cudaMemcpyDeviceToHost));
    DevBuf d_out, d_start;
    CUDA_TRY(dev_alloc(&d_out.p, (size_t)cap * 4));
    CUDA_TRY(dev_alloc(&d_start.p, 4));
    split_path_kernel<<<1, 1>>>(split.sp.t_parent, split.sp.t_lit, split.sp.t_seg_off, split.sp.t_seg_len, split.sp.log,
                                leaf, d_out.as<int32_t>(), cap, d_start.as<int32_t>());
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(&start, d_start.p, 4, cudaMemcpyDeviceToHost));
    if (start < 0) return fail(NDP_E_CUDA, "split path longer than the variable count");
    std::vector<int32_t> path((size_t)(cap - start));
    if (!path.empty())
        CUDA_TRY(cudaMemcpy(path.data(), d_out.as<int32_t>() + start, path.size() * 4, cudaMemcpyDeviceToHost));
    m.insert(m.end(), path.begin(), path.end());
    if (stats) { stats->launches += 1; stats->d2h_bytes += 8 + path.size() * 4; }
    return NDP_OK;
}

int batch_over_pieces(ndp_frontier* f, const IndexRoot& ix, SplitRun& split, int n_gpus, int early_exit,
                      uint8_t* verdict, size_t* winner, int32_t* model, size_t model_cap, size_t* model_len,
                      ndp_stats* stats) {
    const size_t n_sub = f->entries.size(), np = split.n_pieces;
    if (winner) *winner = SIZE_MAX;
    if (model_len) *model_len = 0;
    PieceSet ps;
    int rc = make_piece_set(f, ix, split, ps);
    if (rc) return rc;
    std::vector<uint8_t> pv(std::max(np, (size_t)1), NDP_NOT_RUN);
    std::vector<uint64_t> pn(std::max(np, (size_t)1), 0), pe(std::max(np, (size_t)1), 0);
    std::vector<int32_t> tail(model_cap ? model_cap : 1);
    size_t pwin = SIZE_MAX, tail_len = 0;
    ndp_stats st;
    memset(&st, 0, sizeof st);
    if (np) rc = batch_core(ps.pf.get(), n_gpus, early_exit, pv.data(), &pwin, tail.data(), tail.size(), &tail_len, pn.data(),
                            pe.data(), &st);
    ps.release();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(f->device));
    // fold the pieces back: first piece with a model decides the subproblem; counters as in dfs_on_device
    std::vector<size_t> first_sat(n_sub, SIZE_MAX);
    for (size_t pc = 0; pc < np; ++pc)
        if (pv[pc] == NDP_SAT && first_sat[(size_t)ps.orig[pc]] == SIZE_MAX) first_sat[(size_t)ps.orig[pc]] = pc;
    std::vector<uint8_t> sub_v(n_sub, NDP_UNSAT);
    std::vector<unsigned long long> sub_n(n_sub, 0), sub_e(n_sub, 0);
    for (size_t pc = 0; pc < np; ++pc) {
        const size_t o = (size_t)ps.orig[pc];
        if (first_sat[o] != SIZE_MAX && pc > first_sat[o]) continue;
        sub_n[o] += ps.pre_n[pc] + pn[pc];
        sub_e[o] += ps.pre_e[pc] + pe[pc];
        if (pv[pc] == NDP_NOT_RUN && sub_v[o] == NDP_UNSAT) sub_v[o] = NDP_NOT_RUN;
    }
    size_t win = SIZE_MAX;
    uint64_t tn = 0, te = 0;
    for (size_t o = 0; o < n_sub; ++o) {
        if (first_sat[o] != SIZE_MAX) sub_v[o] = NDP_SAT;
        else if (sub_v[o] == NDP_UNSAT) { sub_n[o] += ps.tail_n[o]; sub_e[o] += ps.tail_e[o]; }
        if (verdict) verdict[o] = sub_v[o];
        tn += sub_n[o]; te += sub_e[o];
        if (sub_v[o] == NDP_SAT && win == SIZE_MAX) win = o;
    }
    if (winner) *winner = win;
    if (win != SIZE_MAX && pwin != first_sat[win])
        return fail(NDP_E_CUDA, "piece winner disagrees with the per-subproblem fold");
    if (win != SIZE_MAX && (model || model_len)) {
        // choices_k ++ the split's path to the winning piece ++ what the device that searched it appended
        std::vector<int32_t> m;
        if ((rc = model_prefix(f, split, win, pwin, m, stats))) return rc;
        m.insert(m.end(), tail.begin(), tail.begin() + tail_len);
        if (model_len) *model_len = m.size();
        if (model) {
            if (m.size() > model_cap) return fail(NDP_E_INVALID, "model buffer too small");
            memcpy(model, m.data(), m.size() * sizeof(int32_t));
        }
    }
    if (stats) {
        stats->nodes += tn;
        stats->clause_evals += te;
        stats->rebuilds += st.rebuilds;
        stats->kernel_ms += st.kernel_ms;
        stats->launches += st.launches;
        stats->h2d_bytes += st.h2d_bytes;
        stats->d2h_bytes += st.d2h_bytes + np * (8 + 16 + 4 + 16) + n_sub * 16;
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
            map_out(f, m.data(), m.size());
            memcpy(model, m.data(), m.size() * sizeof(int32_t));
        }
    }
    return NDP_OK;
}

struct ndp_exit_word {
    unsigned long long* ptr = nullptr;   // device memory: own (cudaMalloc) or another process's, mapped through CUDA IPC
    bool imported = false;
    int device = 0;
};

int ndp_exit_word_create(ndp_exit_word** word) {
    if (!word) return fail(NDP_E_INVALID, "word is null");
    *word = nullptr;
    int rc = check_device();
    if (rc) return rc;
    std::unique_ptr<ndp_exit_word> w(new ndp_exit_word);
    w->device = g_device;
    CUDA_TRY(cudaMalloc(&w->ptr, 8));   // a plain allocation: pool memory cannot be exported over IPC
    const unsigned long long none = ~0ull;
    CUDA_TRY(cudaMemcpy(w->ptr, &none, 8, cudaMemcpyHostToDevice));
    *word = w.release();
    return NDP_OK;
}
int ndp_exit_word_export(const ndp_exit_word* word, uint8_t handle[NDP_EXIT_HANDLE_BYTES]) {
    if (!word || !handle || word->imported) return fail(NDP_E_INVALID, "export needs a word this process created");
    static_assert(sizeof(cudaIpcMemHandle_t) == NDP_EXIT_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, word->ptr));
    memcpy(handle, &h, sizeof h);
    return NDP_OK;
}
int ndp_exit_word_import(const uint8_t handle[NDP_EXIT_HANDLE_BYTES], ndp_exit_word** peer) {
    if (!handle || !peer) return fail(NDP_E_INVALID, "null argument");
    *peer = nullptr;
    int rc = check_device();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    std::unique_ptr<ndp_exit_word> w(new ndp_exit_word);
    w->imported = true;
    w->device = g_device;
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    w->ptr = static_cast<unsigned long long*>(p);
    *peer = w.release();
    return NDP_OK;
}
int ndp_exit_word_reset(ndp_exit_word* word) {
    if (!word || word->imported) return fail(NDP_E_INVALID, "reset needs a word this process created");
    CUDA_TRY(cudaSetDevice(word->device));
    const unsigned long long none = ~0ull;
    CUDA_TRY(cudaMemcpy(word->ptr, &none, 8, cudaMemcpyHostToDevice));
    return NDP_OK;
}
int ndp_exit_word_read(const ndp_exit_word* word, size_t* lowest) {
    if (!word || !lowest) return fail(NDP_E_INVALID, "null argument");
    unsigned long long v = ~0ull;
    CUDA_TRY(cudaMemcpy(&v, word->ptr, 8, cudaMemcpyDeviceToHost));
    *lowest = v == ~0ull ? SIZE_MAX : (size_t)v;
    return NDP_OK;
}
void ndp_exit_word_free(ndp_exit_word* word) {
    if (!word) return;
    if (word->ptr) {
        cudaSetDevice(word->device);
        if (word->imported) cudaIpcCloseMemHandle(word->ptr);
        else cudaFree(word->ptr);
    }
    delete word;
}

int ndp_dfs_shard_shared(ndp_frontier* f, size_t first, size_t stride, int early_exit, ndp_exit_word* own,
                         ndp_exit_word* const* peers, int n_peers, uint8_t* verdict, size_t* winner,
                         int32_t* model, size_t model_cap, size_t* model_len, uint64_t* nodes, uint64_t* evals,
                         ndp_stats* stats) {
    if (!f) return fail(NDP_E_INVALID, "frontier is null");
    if (stride == 0) return fail(NDP_E_INVALID, "stride must be >= 1");
    if (!own || own->imported) return fail(NDP_E_INVALID, "own must be a word this process created");
    if (n_peers < 0 || n_peers > 7 || (n_peers && !peers)) return fail(NDP_E_INVALID, "0..7 peer words");
    int rc = check_device();
    if (rc) return rc;
    if (own->device != f->device) return fail(NDP_E_INVALID, "the exit word lives on another device than the frontier");
    if (stats) memset(stats, 0, sizeof *stats);
    if (model_len) *model_len = 0;
    std::vector<unsigned long long*> all(1, own->ptr);
    for (int i = 0; i < n_peers; ++i) {
        if (!peers[i]) return fail(NDP_E_INVALID, "peer word is null");
        all.push_back(peers[i]->ptr);
    }
    std::vector<int32_t> m;
    size_t win = SIZE_MAX;
    rc = dfs_on_device(f, f->device, first, stride, early_exit, own->ptr, all.data(), (int)all.size(), ~0ull, verdict, &win,
                       &m, nodes, evals, stats);
    if (rc) return rc;
    if (winner) *winner = win;
    if (win != SIZE_MAX) {
        if (model_len) *model_len = m.size();
        if (model && !m.empty()) {
            if (m.size() > model_cap) return fail(NDP_E_INVALID, "model buffer too small");
            map_out(f, m.data(), m.size());
            memcpy(model, m.data(), m.size() * sizeof(int32_t));
        }
    }
    return NDP_OK;
}

// ---- piece-level sharing between processes ---------------------------------------------------
struct ndp_piece_run {
    ndp_frontier* f = nullptr;
    size_t rank = 0, world = 1, n_sub = 0, np = 0;
    bool cut = false;                     // false: the pieces ARE the subproblems (frontier wide enough)
    SplitRun split;                       // owns the split's arenas (cut == true)
    PieceSet ps;
    std::vector<uint8_t> pv;              // [pieces] verdicts of this rank's pieces (others NDP_NOT_RUN / untouched)
    std::vector<uint64_t> pn, pe;         // [pieces] own nodes / evals
    size_t pwin = SIZE_MAX;               // lowest own piece with a model
    std::vector<int32_t> tail;            // the DFS choices of pwin
};

int ndp_dfs_pieces_begin(ndp_frontier* f, size_t rank, size_t world, int early_exit, ndp_exit_word* own,
                         ndp_exit_word* const* peers, int n_peers, ndp_piece_run** run_out, uint32_t* first_model,
                         ndp_stats* stats) {
    if (!f || !run_out) return fail(NDP_E_INVALID, "null argument");
    *run_out = nullptr;
    if (world == 0 || rank >= world) return fail(NDP_E_INVALID, "rank must be < world");
    if (own && own->imported) return fail(NDP_E_INVALID, "own must be a word this process created");
    if (n_peers < 0 || n_peers > 7 || (n_peers && (!peers || !own))) return fail(NDP_E_INVALID, "0..7 peer words (with own)");
    int rc = check_device();
    if (rc) return rc;
    if (own && own->device != f->device) return fail(NDP_E_INVALID, "the exit word lives on another device than the frontier");
    if (stats) memset(stats, 0, sizeof *stats);
    std::unique_ptr<ndp_piece_run> run(new ndp_piece_run);
    run->f = f; run->rank = rank; run->world = world;
    run->n_sub = f->entries.size();
    if (run->n_sub > 0xfffffff0ull) return fail(NDP_E_UNSUPPORTED, "frontier larger than the 32-bit piece index");
    const double t0 = now_ms();
    // the cut: identical on every rank (same frontier, same device model, same world)
    const IndexRoot* ix = nullptr;
    if (engine_uses_index(f) && run->n_sub) {
        ndp_frontier::IndexReplica* rep = nullptr;
        if ((rc = build_index_replica(f, f->device, 0, 1, &ix, &rep))) return rc;
        CUDA_TRY(cudaSetDevice(f->device));
        IndexLaunchPlan iplan;
        if ((rc = plan_index(f->device, *ix, f->max_var, iplan))) return rc;
        // The job-wide cut is made by EVERY rank (replicated work), so it is coarse: enough pieces for the
        // static deal piece -> rank to balance statistically (>= 512 per rank: the subproblems of a granulated
        // multiplier CNF are nearly uniform -- reference node counts, tests/golden/big_*.npz: cv 0.001-0.03 on the
        // 48-bit frontiers, 0.2 on the 32-bit ones -- so 512 per rank already balance to well under 1 %), and only where cutting pays
        // (split_target's test: shard starved or subproblems still deep).  Each rank then refines its own
        // share down to pieces-per-warp granularity inside dfs_on_device -- in parallel, on its own data.
        const size_t n_warps = (size_t)iplan.searchers() * world;
        size_t target = split_target(run->n_sub, (int)std::min<size_t>(n_warps, 1u << 24), ix->n_root, rep->mean_alive, iplan.groups);
        if (g_split < 0 && !getenv("NDP_SPLIT")) {            // automatic policy
            const size_t coarse = (size_t)512 * world;
            target = (world > 1 && target > run->n_sub && coarse > run->n_sub) ? coarse : 0;
        }
        if (target > run->n_sub) {
            if ((rc = run_split(f->device, *ix, *rep, target, run->split, stats))) return rc;
            run->cut = run->split.active;
        }
    }
    ndp_frontier* pf = f;
    if (run->cut) {
        if ((rc = make_piece_set(f, *ix, run->split, run->ps))) return rc;
        pf = run->ps.pf.get();
        run->np = run->split.n_pieces;
    } else {
        run->np = run->n_sub;
    }
    const double split_ms = now_ms() - t0;
    const size_t np1 = std::max(run->np, (size_t)1);
    run->pv.assign(np1, NDP_NOT_RUN);
    run->pn.assign(np1, 0);
    run->pe.assign(np1, 0);
    std::vector<unsigned long long*> all;
    if (own) {
        all.push_back(own->ptr);
        for (int i = 0; i < n_peers; ++i) {
            if (!peers[i]) return fail(NDP_E_INVALID, "peer word is null");
            all.push_back(peers[i]->ptr);
        }
    }
    ndp_stats st;
    memset(&st, 0, sizeof st);
    if (run->np) {
        rc = dfs_on_device(pf, f->device, rank, world, early_exit, own ? own->ptr : nullptr, own ? all.data() : nullptr,
                           (int)all.size(), ~0ull, run->pv.data(), &run->pwin, &run->tail, run->pn.data(), run->pe.data(), &st);
        if (rc) return rc;
    }
    if (stats) {
        stats->kernel_ms += st.kernel_ms + (run->cut ? split_ms : 0.0);   // the cut is part of the search
        stats->launches += st.launches; stats->h2d_bytes += st.h2d_bytes; stats->d2h_bytes += st.d2h_bytes;
        stats->rebuilds += st.rebuilds;
    }
    if (first_model) {
        for (size_t o = 0; o < run->n_sub; ++o) first_model[o] = 0xffffffffu;
        for (size_t pc = rank; pc < run->np; pc += world)
            if (run->pv[pc] == NDP_SAT) {
                const size_t o = run->cut ? (size_t)run->ps.orig[pc] : pc;
                if (first_model[o] == 0xffffffffu) first_model[o] = (uint32_t)pc;
            }
    }
    *run_out = run.release();
    return NDP_OK;
}

int ndp_dfs_pieces_fold(ndp_piece_run* run, const uint32_t* first_model, uint8_t* verdict, uint64_t* nodes,
                        uint64_t* evals, size_t* winner, int32_t* model, size_t model_cap, size_t* model_len,
                        ndp_stats* stats) {
    if (!run || !first_model) return fail(NDP_E_INVALID, "null argument");
    if (model_len) *model_len = 0;
    if (winner) *winner = SIZE_MAX;
    const size_t n_sub = run->n_sub, np = run->np, rank = run->rank, world = run->world;
    std::vector<uint8_t> sub_v(n_sub, NDP_UNSAT);
    std::vector<uint64_t> sub_n(n_sub, 0), sub_e(n_sub, 0);
    for (size_t pc = rank; pc < np; pc += world) {
        const size_t o = run->cut ? (size_t)run->ps.orig[pc] : pc;
        if (first_model[o] != 0xffffffffu && pc > (size_t)first_model[o]) continue;   // the reference stops at the first model
        sub_n[o] += run->pn[pc] + (run->cut ? run->ps.pre_n[pc] : 0ull);
        sub_e[o] += run->pe[pc] + (run->cut ? run->ps.pre_e[pc] : 0ull);
        if (run->pv[pc] == NDP_NOT_RUN) sub_v[o] = NDP_NOT_RUN;
    }
    size_t win = SIZE_MAX;
    uint64_t tn = 0, te = 0;
    for (size_t o = 0; o < n_sub; ++o) {
        if (first_model[o] != 0xffffffffu) {
            sub_v[o] = NDP_SAT;                                   // on every rank: the word is job-wide
            if (win == SIZE_MAX) win = o;
        } else if (run->cut && o % world == rank) {               // dead ends after the last piece: one rank adds them
            sub_n[o] += run->ps.tail_n[o];
            sub_e[o] += run->ps.tail_e[o];
        }
        if (verdict) verdict[o] = sub_v[o];
        if (nodes) nodes[o] = sub_n[o];
        if (evals) evals[o] = sub_e[o];
        tn += sub_n[o]; te += sub_e[o];
    }
    if (stats) { stats->nodes = tn; stats->clause_evals = te; }
    if (winner) *winner = win;
    if (win != SIZE_MAX && (size_t)first_model[win] % world == rank && (model || model_len)) {
        // this rank searched the winning piece: choices_k ++ the cut's path to it ++ its DFS choices
        const size_t pwin = (size_t)first_model[win];
        if (run->pwin != pwin) return fail(NDP_E_CUDA, "piece winner disagrees with the per-subproblem fold");
        std::vector<int32_t> m;   // (without a cut dfs_on_device already returned choices_k ++ DFS choices)
        int rc;
        if (run->cut && (rc = model_prefix(run->f, run->split, win, pwin, m, stats))) return rc;
        m.insert(m.end(), run->tail.begin(), run->tail.end());
        if (model_len) *model_len = m.size();
        if (model) {
            if (m.size() > model_cap) return fail(NDP_E_INVALID, "model buffer too small");
            map_out(run->f, m.data(), m.size());
            memcpy(model, m.data(), m.size() * sizeof(int32_t));
        }
    }
    return NDP_OK;
}

void ndp_dfs_pieces_end(ndp_piece_run* run) {
    if (!run) return;
    cudaSetDevice(run->f->device);
    run->ps.release();
    delete run;
}

int ndp_dfs_progress(size_t* units_pulled, size_t* units_total, int* devices_active) {
    unsigned long long pulled = 0, total = 0;
    int active = 0;
    for (ProgressSlot& ps : g_progress) {
        if (!ps.active.load() || !ps.pulled) continue;
        const unsigned long long t = ps.total.load();
        unsigned long long q = (unsigned long long)*ps.pulled + 1ull;   // the word holds the LAST index handed out
        if (*ps.pulled == 0u && t == 0) q = 0;
        pulled += std::min(q, t);
        total += t;
        ++active;
    }
    if (units_pulled) *units_pulled = (size_t)pulled;
    if (units_total) *units_total = (size_t)total;
    if (devices_active) *devices_active = active;
    return NDP_OK;
}

int ndp_dfs_batch(ndp_frontier* f, int n_gpus, int early_exit, uint8_t* verdict, size_t* winner,
                  int32_t* model, size_t model_cap, size_t* model_len, ndp_stats* stats) {
    if (!f) return fail(NDP_E_INVALID, "frontier is null");
    if (n_gpus <= 1)
        return ndp_dfs_shard(f, 0, 1, early_exit, nullptr, verdict, winner, model, model_cap, model_len, nullptr,
                             nullptr, stats);
    int rc = check_device();
    if (rc) return rc;
    // A frontier too narrow for n_gpus devices is first cut into DFS-order pieces on its home device;
    // the PIECES are then what the devices share out (each may cut its own pieces further).
    if (engine_uses_index(f) && !f->entries.empty()) {
        const IndexRoot* ix = nullptr;
        ndp_frontier::IndexReplica* rep = nullptr;
        if ((rc = build_index_replica(f, f->device, 0, 1, &ix, &rep))) return rc;
        CUDA_TRY(cudaSetDevice(f->device));
        IndexLaunchPlan iplan;
        if ((rc = plan_index(f->device, *ix, f->max_var, iplan))) return rc;
        const int n_warps = iplan.searchers();
        double alive = 0;
        for (const Entry& e : f->entries) alive += (double)e.n;
        alive /= (double)f->entries.size();
        // Same rule as ndp_dfs_pieces_begin: the job-wide cut is made on ONE device and its pieces travel to the others, so it
        // is only made where the static deal of whole subproblems could not balance the devices (fewer than 512 each), and it is
        // coarse (512 pieces per device); every device then refines its own share in parallel.
        const bool worth = 4 * f->entries.size() < (size_t)n_warps / iplan.groups * n_gpus || alive > 0.5 * (double)ix->n_root;
        size_t target = worth ? split_target(f->entries.size(), n_warps * n_gpus, 0, alive, iplan.groups) : 0;
        if (g_split < 0 && !getenv("NDP_SPLIT")) {
            const size_t coarse = (size_t)512 * n_gpus;
            target = (target > f->entries.size() && coarse > f->entries.size()) ? coarse : 0;
        }
        if (target > f->entries.size()) {
            if (stats) memset(stats, 0, sizeof *stats);
            SplitRun split;
            const double t_split = now_ms();
            if ((rc = run_split(f->device, *ix, *rep, target, split, stats))) return rc;
            if (stats) stats->kernel_ms += now_ms() - t_split;   // run_split returns after the kernel (it reads its state back)
            if (split.active) {
                size_t len = 0;
                rc = batch_over_pieces(f, *ix, split, n_gpus, early_exit, verdict, winner, model, model_cap, &len, stats);
                if (!rc && model) map_out(f, model, len);
                if (model_len) *model_len = len;
                return rc;
            }
        }
    }
    size_t len = 0;
    rc = batch_core(f, n_gpus, early_exit, verdict, winner, model, model_cap, &len, nullptr, nullptr, stats);
    if (!rc && model) map_out(f, model, len);
    if (model_len) *model_len = len;
    return rc;
}

}  // extern "C"
