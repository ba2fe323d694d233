// ndp_kernels.cuh -- the two hot kernels: one BFS granulation level, and the persistent
// warp-per-subproblem DFS.  See DESIGN.md for the data layout and the rooflines.
#pragma once
#include "ndp_clause.cuh"

namespace ndp {

// ------------------------------------------------------------------------------------------
// BFS level (Satisfy_iterative_BFS loop body, NDP-5_6_7.cpp:898-925, for every node of one
// level at once).  One warp per node: it reads the node's clause slot once per open branch and
// writes the children LA -> slot 2j, RA -> slot 2j+1 of the other arena buffer (LA before RA,
// NDP:904-925).  A child is "open" iff non-empty and conflict-free -- the push conditions at
// NDP:908 and NDP:919.  The host applies the exact -d/-q cut to the per-child metadata.
struct BfsNode {       // one queue element of the current level
    uint32_t slot;     // slot in the source buffer
    int32_t n;         // clauses in it
    int32_t kind;      // SetInfo.kind of the node's set
    int32_t lit;       // SetInfo.lit
};
struct BfsChild {
    int32_t n;         // emitted clauses; -1 = branch not open (conflict, empty, or not attempted)
    int32_t kind;
    int32_t lit;
    int32_t empty;     // 1 iff the branch emitted no clause at all (a dropped solution leaf)
};

__global__ void __launch_bounds__(256) bfs_expand_kernel(const pclause* __restrict__ src, pclause* __restrict__ dst,
                                                          size_t stride, const BfsNode* __restrict__ nodes, int m,
                                                          BfsChild* __restrict__ children) {
    const unsigned lane = threadIdx.x & 31u;
    const int warps_per_block = blockDim.x >> 5;
    for (int j = blockIdx.x * warps_per_block + (threadIdx.x >> 5); j < m; j += gridDim.x * warps_per_block) {
        BfsNode nd = nodes[j];
        const pclause* A = src + (size_t)nd.slot * stride;
        BfsChild ch[2];
        ch[0].n = ch[1].n = -1;
        ch[0].kind = ch[1].kind = K_EMPTY;
        ch[0].lit = ch[1].lit = 0;
        ch[0].empty = ch[1].empty = 0;
        const int var = nd.kind == K_UNIT ? (nd.lit < 0 ? -nd.lit : nd.lit) : nd.lit;
        if (nd.n > 0 && var != 0) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int t = b == 0 ? var : -var;   // b=0: LA (i := true), b=1: RA (i := false)
                // the branch that falsifies a unit clause always emits {0,0,0}: never open
                if (nd.kind == K_UNIT && t != nd.lit) continue;
                SetInfo s = warp_apply_copy(A, nd.n, dst + ((size_t)2 * j + b) * stride, t, lane);
                if (s.n == 0) ch[b].empty = 1;
                if (s.n > 0 && !s.conflict) { ch[b].n = s.n; ch[b].kind = s.kind; ch[b].lit = s.lit; }
            }
        }
        if (lane == 0) { children[2 * j] = ch[0]; children[2 * j + 1] = ch[1]; }
    }
}

// After a level has been expanded: turn the per-child metadata into the next level's node
// list, in queue order (children of node 0 first, LA before RA -- NDP:904-925), append the
// pushed children to the choice tree, and reduce the level's counters.  One CTA walks the
// 2m children with a block-wide exclusive scan of the "open" flags (warp ballots + a 32-entry
// shared array), so the frontier is compacted without the host ever seeing it.
struct BfsLevelTotals {
    unsigned m_next;             // open children = size of the next level
    unsigned expanded;           // nodes that count as an iteration (choice() != 0, NDP:901-902, 926)
    unsigned long long evals;    // sum of |A| over those nodes
    unsigned max_n;              // largest child
    unsigned stop_q;             // 1 iff the -q test "queue size >= Q before a pop" (NDP:894) fires inside this level
    unsigned pad[2];
};

__global__ void __launch_bounds__(1024) bfs_compact_kernel(const BfsNode* __restrict__ nodes,
                                                           const int32_t* __restrict__ node_tree,
                                                           const BfsChild* __restrict__ children, int m,
                                                           BfsNode* __restrict__ next_nodes,
                                                           int32_t* __restrict__ next_tree,
                                                           int32_t* __restrict__ tree_parent,
                                                           int32_t* __restrict__ tree_lit, int tree_base, int Q,
                                                           BfsLevelTotals* __restrict__ totals) {
    __shared__ unsigned warp_count[32];
    __shared__ unsigned stop_flag;
    if (threadIdx.x == 0) stop_flag = 0u;
    __shared__ unsigned long long red_evals[32];
    __shared__ unsigned red_exp[32], red_max[32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lt = lanemask_lt();
    unsigned running = 0;
    unsigned long long evals = 0;
    unsigned expanded = 0, max_n = 0;
    for (int base = 0; base < 2 * m; base += 1024) {
        const int idx = base + (int)tid;
        BfsChild ch;
        ch.n = -1;
        if (idx < 2 * m) ch = children[idx];
        const bool open = ch.n >= 0;
        const unsigned bal = __ballot_sync(kFullMask, open);
        if (lane == 0) warp_count[warp] = __popc(bal);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const unsigned c = warp_count[w];
            if (w < (int)warp) before += c;
            total += c;
        }
        const unsigned pos = running + before + __popc(bal & lt);   // open children before this one
        // queue size just before node j is popped: the unexpanded rest of the level plus the
        // children pushed so far
        if (Q > 0 && idx < 2 * m && (idx & 1) == 0 && (unsigned)(m - (idx >> 1)) + pos >= (unsigned)Q) stop_flag = 1u;
        if (open) {
            const int j = idx >> 1;
            const BfsNode nd = nodes[j];
            const int var = nd.kind == K_UNIT ? (nd.lit < 0 ? -nd.lit : nd.lit) : nd.lit;
            BfsNode nn;
            nn.slot = (uint32_t)idx;
            nn.n = ch.n;
            nn.kind = ch.kind;
            nn.lit = ch.lit;
            next_nodes[pos] = nn;
            next_tree[pos] = tree_base + (int)pos;
            tree_parent[tree_base + pos] = node_tree[j];
            tree_lit[tree_base + pos] = (idx & 1) ? -var : var;
            max_n = max(max_n, (unsigned)ch.n);
        }
        running += total;
        __syncthreads();
    }
    for (int j = (int)tid; j < m; j += 1024) {
        const BfsNode nd = nodes[j];
        const int var = nd.kind == K_UNIT ? (nd.lit < 0 ? -nd.lit : nd.lit) : nd.lit;
        if (nd.n > 0 && var != 0) { evals += (unsigned long long)nd.n; ++expanded; }
    }
    for (int off = 16; off > 0; off >>= 1) evals += __shfl_down_sync(kFullMask, evals, off);
    expanded = __reduce_add_sync(kFullMask, expanded);
    max_n = __reduce_max_sync(kFullMask, max_n);
    if (lane == 0) { red_evals[warp] = evals; red_exp[warp] = expanded; red_max[warp] = max_n; }
    __syncthreads();
    if (tid == 0) {
        BfsLevelTotals t;
        t.m_next = running; t.expanded = 0; t.evals = 0; t.max_n = 0;
        t.stop_q = stop_flag;
        t.pad[0] = t.pad[1] = 0;
        for (int w = 0; w < 32; ++w) { t.evals += red_evals[w]; t.expanded += red_exp[w]; t.max_n = max(t.max_n, red_max[w]); }
        *totals = t;
    }
}

// ------------------------------------------------------------------------------------------
// DFS (Satisfy_iterative(A, true), NDP-5_6_7.cpp:784-878).  Persistent grid; every warp pulls
// the next subproblem of its shard from an atomic counter and runs the whole search:
//   W      working clause set, compacted in place after every assignment (shared memory,
//          or a per-warp global scratch slot when a subproblem does not fit);
//   trail  the choices made so far in decision order (the reference's `choices` vector);
//   pend   one bit per trail position: "the LA sibling of this RA decision is still to be
//          explored" -- the only stack the search needs, because a sibling is re-materialised
//          from the subproblem's base slot under the trail's assignment when it is popped;
//   table  2-bit assignment table used by that re-materialisation.
// Order of exploration is the reference's: RA (i := false) before LA, an empty LA wins at once.
struct DfsParams {
    // per-subproblem arrays are indexed by the LOCAL shard index q
    const pclause* const* sub_base;  // [n_local] device pointer to each subproblem's base slot
    const int32_t* sub_n;            // [n_local] clauses in it
    unsigned long long first, stride, n_local;  // shard: global index g = first + q*stride
    unsigned int* next;              // work counter
    unsigned long long* best;        // lowest SAT global index seen (ULLONG_MAX = none)
    unsigned long long* peer_best[8];
    int n_peers;
    int early_exit;
    uint8_t* verdict;                // [n_local]
    unsigned long long* nodes;       // [n_local]
    unsigned long long* evals;       // [n_local]
    unsigned int* rebuilds;          // [n_local]
    unsigned long long* model_sub;   // [n_warps] which subproblem (global index) the warp's model slot holds
    int32_t* model_len;              // [n_warps]
    int32_t* models;                 // [n_warps][model_stride]
    int model_stride;
    pclause* gscratch;               // [n_warps][w_cap] when W lives in global memory
    int w_cap;                       // clause capacity of W
    int trail_cap;                   // = max variable
    int warp_bytes, off_trail, off_table, off_pend, table_words, pend_words;
};

template <bool kSmemW>
__global__ void dfs_kernel(const __grid_constant__ DfsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned char* mine = smem_raw + (size_t)warp * p.warp_bytes;
    pclause* W = kSmemW ? reinterpret_cast<pclause*>(mine) : p.gscratch + (size_t)gwarp * p.w_cap;
    short* trail = reinterpret_cast<short*>(mine + p.off_trail);
    unsigned* table = reinterpret_cast<unsigned*>(mine + p.off_table);
    unsigned* pend = reinterpret_cast<unsigned*>(mine + p.off_pend);

    for (;;) {
        unsigned q = 0;
        if (lane == 0) q = atomicAdd(p.next, 1u);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= p.n_local) break;
        const unsigned long long g = p.first + (unsigned long long)q * p.stride;
        if (p.early_exit && *(volatile unsigned long long*)p.best < g) {
            if (lane == 0) { p.verdict[q] = 2; p.nodes[q] = 0; p.evals[q] = 0; p.rebuilds[q] = 0; }
            continue;
        }
        const pclause* base = p.sub_base[q];
        const int nb = p.sub_n[q];
        for (int w = lane; w < p.pend_words; w += 32) pend[w] = 0u;
        int ntrail = 0;
        SetInfo cur = warp_rebuild<false>(base, nb, W, nullptr, lane);
        int n_buf = cur.n;                 // records in W (live + tombstones)
        bool sticky = cur.conflict != 0;   // a {0,0,0} clause can never leave the set
        __syncwarp();
        unsigned long long nodes = 0, evals = 0;
        unsigned rebuilds = 0;
        int verdict = 0;  // UNSAT unless a model turns up
        for (;;) {
            ++nodes;  // one Satisfy_iterative loop iteration (NDP:801-806)
            if (p.early_exit && (nodes & 31u) == 0 && *(volatile unsigned long long*)p.best < g) { verdict = 2; break; }
            const int var = cur.kind == K_UNIT ? (cur.lit < 0 ? -cur.lit : cur.lit) : cur.lit;
            if (cur.n == 0 || var == 0) { verdict = 1; break; }  // choice()==0: record choices (NDP:817-823)
            evals += (unsigned long long)cur.n;                 // one ResolutionStepWithConflict (NDP:831)
            // Which branch is explored next (t), and is its LA sibling left pending?
            int t = 0;
            bool pending = false;
            if (cur.kind == K_UNIT) {
                // i = |x|: the branch that makes x false emits {0,0,0} (conflict, never pushed, NDP:837/857)
                t = cur.lit;
            } else {
                SetInfo la = warp_pass_lazy<false>(W, n_buf, cur.n, var, sticky, lane);
                if (la.n == 0) {                                 // LA empty wins before RA is looked at (NDP:843-850)
                    if (lane == 0) trail[ntrail] = (short)var;
                    ++ntrail; verdict = 1; break;
                }
                SetInfo ra = warp_pass_lazy<false>(W, n_buf, cur.n, -var, sticky, lane);
                if (ra.n == 0) {                                 // NDP:863-870
                    if (lane == 0) trail[ntrail] = (short)(-var);
                    ++ntrail; verdict = 1; break;
                }
                if (!ra.conflict) { t = -var; pending = !la.conflict; }  // RA is on top of the stack: explored first
                else if (!la.conflict) t = var;
            }
            bool dead = t == 0;
            if (!dead) {
                if (lane == 0) {
                    if (pending) pend[ntrail >> 5] |= 1u << (ntrail & 31);
                    trail[ntrail] = (short)t;
                }
                ++ntrail;
                SetInfo r = warp_pass_lazy<true>(W, n_buf, cur.n, t, sticky, lane);
                __syncwarp();
                if (r.n == 0) { verdict = 1; break; }           // empty branch => model (NDP:843-852 / 863-872)
                if (r.conflict) dead = true;
                else {
                    cur = r;
                    if (n_buf - cur.n > (n_buf >> 3) + 32) {     // squeeze out tombstones now and then
                        n_buf = warp_compact(W, n_buf, lane);
                        __syncwarp();
                    }
                    continue;
                }
            }
            // Backtrack: pop to the deepest decision whose LA sibling is still pending.
            __syncwarp();
            int pos = -1;
            for (int w = (ntrail - 1) >> 5; w >= 0 && ntrail > 0; --w) {
                unsigned bits = pend[w];
                if (w == ((ntrail - 1) >> 5)) {
                    int top = (ntrail - 1) & 31;
                    if (top != 31) bits &= (2u << top) - 1u;
                }
                if (bits) { pos = w * 32 + 31 - __clz(bits); break; }
            }
            if (pos < 0) break;  // stack empty: UNSAT
            __syncwarp();
            if (lane == 0) {
                pend[pos >> 5] &= ~(1u << (pos & 31));
                trail[pos] = (short)(-trail[pos]);  // -i -> +i : now the LA sibling
            }
            ntrail = pos + 1;
            for (int w = lane; w < p.table_words; w += 32) table[w] = 0u;
            __syncwarp();
            for (int j = lane; j < ntrail; j += 32) {
                int l = trail[j];
                unsigned v = (unsigned)(l < 0 ? -l : l);
                atomicOr(&table[v >> 4], (l > 0 ? 1u : 2u) << ((v & 15u) * 2u));
            }
            __syncwarp();
            cur = warp_rebuild<true>(base, nb, W, table, lane);
            n_buf = cur.n;
            __syncwarp();
            ++rebuilds;
        }
        if (verdict == 1) {
            unsigned long long old = ~0ull;
            if (lane == 0) {
                old = atomicMin(p.best, g);
                if (g < old)
                    {
#pragma unroll
                        for (int d = 0; d < 8; ++d)   // constant indices: the parameter block stays in constant memory
                            if (d < p.n_peers) atomicMin_system(p.peer_best[d], g);
                    }
            }
            old = __shfl_sync(kFullMask, old, 0);
            if (g < old) {  // this warp now holds the lowest-index model: park it in its slot
                __syncwarp();
                int32_t* out = p.models + (size_t)gwarp * p.model_stride;
                for (int j = lane; j < ntrail; j += 32) out[j] = (int32_t)trail[j];
                if (lane == 0) { p.model_sub[gwarp] = g; p.model_len[gwarp] = ntrail; }
            }
        }
        if (lane == 0) {
            p.verdict[q] = (uint8_t)verdict;
            p.nodes[q] = nodes;
            p.evals[q] = evals;
            p.rebuilds[q] = rebuilds;
        }
        __syncwarp();
    }
}

}  // namespace ndp
