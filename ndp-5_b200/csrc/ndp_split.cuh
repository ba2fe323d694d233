// ndp_split.cuh -- DFS-order split of a narrow shard into pieces (index engine).
//
// Satisfy_iterative (NDP-5_6_7.cpp:784-878) is sequential inside one subproblem, and a frontier
// can be far narrower than the GPU: -d counts node expansions (NDP:926,935), so BASELINE configs
// 2 and 4 (-d 12, -d 20) yield ONE subproblem, and a 1-of-8 shard of a modest -q frontier holds
// fewer subproblems than the device has warps.  The split turns each subproblem into an ordered
// list of PIECES -- subtrees of its search tree, listed in the order the reference's LIFO stack
// would reach them (RA before LA: NDP:857-862 is pushed last and popped first) -- so that
//
//   * the first model in piece order is the model the reference finds (lowest piece with a model
//     = first model in DFS pre-order), and the subproblem's verdict is "some piece has a model";
//   * node and clause-evaluation counts stay exact: every node the split itself visits (the
//     interior of the tree above the pieces) is charged to the first piece that follows it in
//     pre-order ("pre" counts), so  nodes(subproblem) = sum over pieces up to and including the
//     winning piece of (pre + own)  -- exactly the nodes the reference pops before it stops --
//     and for an UNSAT subproblem the sum over all pieces plus the "tail" (dead ends after the
//     last piece).
//
// One split level = every piece descends its chain of forced moves (unit clauses, NDP:756-766,
// and decisions with one conflicting child) until it reaches a decision whose two children are
// both open, and is replaced by those two children, RA first.  99.6 % of the reference's nodes
// are such forced moves, so a level roughly doubles the piece count for the cost of ~60 nodes per
// warp.  A descent that hits a model becomes a LEAF piece (no node left to visit), a dead end
// disappears and forwards its counts to the next piece in order.  The whole loop is ONE
// cooperative kernel (expand | grid.sync | ordered compaction by CTA 0 | grid.sync).
//
// Choices made inside the split are logged (a parent-pointer tree of chain segments) so the
// winner's model is  choices_k ++ split path ++ DFS trail  (NDP:1430-1431).
#pragma once
#include "ndp_index.cuh"

namespace ndp {

enum : int { SPLIT_PASS = 0, SPLIT_ADVANCED = 1, SPLIT_LEAF = 2, SPLIT_BRANCH = 3, SPLIT_DEAD = 4 };
enum : int { SEXIT_NONE = 0, SEXIT_TARGET = 1, SEXIT_EMPTY = 2, SEXIT_STALLED = 3, SEXIT_CAPACITY = 4, SEXIT_LEVELS = 5 };

struct SplitParent {               // what expanding piece j produced
    int outcome;
    int var;                       // SPLIT_BRANCH: the decision variable
    unsigned chain_nodes;          // nodes the descent visited
    unsigned fwd;                  // SPLIT_DEAD: index (next level) of the first piece after it
    unsigned long long chain_evals;
};

struct SplitLevel {                // piece arrays of one level (two sides, ping-pong)
    const unsigned** ptr;          // state slot (nullptr for a LEAF)
    BfsNode* meta;                 // n / kind / lit (slot unused)
    int32_t* orig;                 // local subproblem index
    int32_t* tnode;                // node in the split tree
    unsigned long long* pre_nodes; // interior nodes charged to this piece
    unsigned long long* pre_evals;
};

struct SplitState {
    int m, cur, levels_done, exit_reason, progressed, pad;
    unsigned long long tree_count, log_count, log_cursor;
    unsigned long long cap_pieces, cap_tree, cap_log;
};

struct SplitParams {
    IndexRoot ix;
    SplitLevel lv[2];
    unsigned* sbuf[2];             // child slots: a level reading side c writes sbuf[c ^ 1][2j + b]
    BfsChild* children;            // [cap_pieces]   (.empty: bit 0 = piece has no slot)
    SplitParent* parents;          // [cap_pieces / 2]
    int32_t* t_parent;             // split tree: parent node (-1 for a subproblem's root)
    int32_t* t_lit;                //   branch literal that created the node (0: none)
    unsigned* t_seg_off;           //   chain segment logged when the node's piece was expanded
    unsigned* t_seg_len;
    short* log;
    unsigned long long* tail_nodes;   // [subproblems] dead ends after the subproblem's last piece
    unsigned long long* tail_evals;
    SplitState* st;
    int chain_cap, max_levels;
    unsigned target;
    int warp_bytes, off_chain;     // shared memory of one warp: state slot, then the chain
};

// Level 0: the shard's subproblems are the pieces.
__global__ void split_init_kernel(const SplitParams p, const unsigned* const* slot_ptr, const BfsNode* meta, int n_sub) {
    const SplitLevel& L = p.lv[0];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_sub; j += gridDim.x * blockDim.x) {
        L.ptr[j] = slot_ptr[j];
        L.meta[j] = meta[j];
        L.orig[j] = j;
        L.tnode[j] = j;
        L.pre_nodes[j] = 0ull;
        L.pre_evals[j] = 0ull;
        p.t_parent[j] = -1;
        p.t_lit[j] = 0;
        p.t_seg_off[j] = 0u;
        p.t_seg_len[j] = 0u;
        p.tail_nodes[j] = 0ull;
        p.tail_evals[j] = 0ull;
    }
}

template <bool kPlain>
__device__ __forceinline__ void split_expand_piece(const SplitParams& p, int cur, int j, unsigned* S, short* chain,
                                                   unsigned lane) {
    const IndexRoot& ix = p.ix;
    const SplitLevel& L = p.lv[cur];
    BfsNode nd;
    {
        const int4 raw = __ldcg(reinterpret_cast<const int4*>(&L.meta[j]));
        nd.slot = (uint32_t)raw.x; nd.n = raw.y; nd.kind = raw.z; nd.lit = raw.w;
    }
    const unsigned* src = reinterpret_cast<const unsigned*>(__ldcg(reinterpret_cast<const unsigned long long*>(&L.ptr[j])));
    unsigned* dst = p.sbuf[cur ^ 1];
    unsigned* d0 = dst + (size_t)2 * j * ix.slot_words;
    unsigned* d1 = d0 + ix.slot_words;
    BfsChild ch[2];
    ch[0].n = ch[1].n = -1;
    ch[0].kind = ch[1].kind = K_EMPTY;
    ch[0].lit = ch[1].lit = 0;
    ch[0].empty = ch[1].empty = 0;
    SplitParent par;
    par.outcome = SPLIT_PASS; par.var = 0; par.chain_nodes = 0u; par.fwd = 0u; par.chain_evals = 0ull;
    const int var0 = nd.kind == K_UNIT ? (nd.lit < 0 ? -nd.lit : nd.lit) : nd.lit;
    if (nd.kind == K_LEAF) {
        ch[0].n = 0; ch[0].kind = K_LEAF; ch[0].empty = 1;
    } else if (nd.n <= 0 || var0 == 0) {
        // choice() == 0 at the subproblem's own root (NDP:817-823): the DFS kernel resolves it in one node
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(d0);
        for (int i = lane; i < ix.slot_words / 4; i += 32) d4[i] = __ldcg(&s4[i]);
        ch[0].n = nd.n; ch[0].kind = nd.kind; ch[0].lit = nd.lit;
    } else {
        unsigned* table = S + ix.table_off;
        __syncwarp();
        slot_load_cg(S, src, ix.slot_words, lane);
        __syncwarp();
        int n_live = nd.n, kind = nd.kind, lit = nd.lit;
        unsigned cn = 0;
        unsigned long long ce = 0;
        int nchain = 0, outcome = SPLIT_ADVANCED;
        for (;;) {
            const int v = kind == K_UNIT ? (lit < 0 ? -lit : lit) : lit;
            if ((int)cn >= p.chain_cap || v == 0) { outcome = SPLIT_ADVANCED; break; }   // carried on as it is
            ++cn;                                   // one Satisfy_iterative loop iteration (NDP:801-806)
            ce += (unsigned long long)n_live;       // one ResolutionStepWithConflict (NDP:831)
            int t;
            if (kind == K_UNIT) {
                t = lit;                            // the other branch falsifies the unit clause
            } else {
                bool la_conf, ra_conf;
                int la_n, ra_n;
                index_classify(ix, S, n_live, v, la_n, la_conf, ra_n, ra_conf, lane);
                if (la_n == 0) {                    // LA empty wins first (NDP:843-850)
                    if (lane == 0) chain[nchain] = (short)v;
                    ++nchain; outcome = SPLIT_LEAF; break;
                }
                if (ra_n == 0) {                    // NDP:863-870
                    if (lane == 0) chain[nchain] = (short)(-v);
                    ++nchain; outcome = SPLIT_LEAF; break;
                }
                if (!la_conf && !ra_conf) { outcome = SPLIT_BRANCH; par.var = v; break; }
                t = !ra_conf ? -v : (!la_conf ? v : 0);
            }
            if (t == 0) { outcome = SPLIT_DEAD; break; }
            if (lane == 0) { chain[nchain] = (short)t; table_set(table, t); }
            ++nchain;
            bool conflict;
            n_live = index_assign<true, kPlain>(ix, S, n_live, t, conflict, lane);
            if (n_live == 0) { outcome = SPLIT_LEAF; break; }        // NDP:843-852 / 863-872
            if (conflict) { outcome = SPLIT_DEAD; break; }
            index_choice<kPlain>(ix, S, table, n_live, lane, kind, lit);
        }
        par.outcome = outcome;
        par.chain_nodes = cn;
        par.chain_evals = ce;
        __syncwarp();
        if (nchain > 0) {      // log the chain as this piece's segment of the split tree
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(&p.st->log_cursor, (unsigned long long)nchain);
            off = __shfl_sync(kFullMask, off, 0);
            for (int i = lane; i < nchain; i += 32) p.log[off + i] = chain[i];
            if (lane == 0) {
                const int32_t tn = __ldcg(&L.tnode[j]);
                p.t_seg_off[tn] = (unsigned)off;
                p.t_seg_len[tn] = (unsigned)nchain;
            }
        }
        if (outcome == SPLIT_ADVANCED) {
            ch[0].n = n_live; ch[0].kind = kind; ch[0].lit = lit;
            slot_copy(d0, S, ix.slot_words, lane);
        } else if (outcome == SPLIT_LEAF) {
            ch[0].n = 0; ch[0].kind = K_LEAF; ch[0].empty = 1;
        } else if (outcome == SPLIT_BRANCH) {
            // children in the order the reference explores them: RA (i := false) first
            const int v = par.var;
            slot_copy(d1, S, ix.slot_words, lane);          // the state before the branch, parked in LA's slot
            __syncwarp();
            bool conflict;
            if (lane == 0) table_set(table, -v);
            int n_c = index_assign<true, kPlain>(ix, S, n_live, -v, conflict, lane);
            int k_c, l_c;
            index_choice<kPlain>(ix, S, table, n_c, lane, k_c, l_c);
            ch[0].n = n_c; ch[0].kind = k_c; ch[0].lit = l_c;
            slot_copy(d0, S, ix.slot_words, lane);
            __syncwarp();
            slot_load_cg(S, d1, ix.slot_words, lane);
            __syncwarp();
            if (lane == 0) table_set(table, v);
            n_c = index_assign<true, kPlain>(ix, S, n_live, v, conflict, lane);
            index_choice<kPlain>(ix, S, table, n_c, lane, k_c, l_c);
            ch[1].n = n_c; ch[1].kind = k_c; ch[1].lit = l_c;
            slot_copy(d1, S, ix.slot_words, lane);
        }
    }
    if (lane == 0) {
        p.children[2 * j] = ch[0];
        p.children[2 * j + 1] = ch[1];
        p.parents[j] = par;
    }
}

template <bool kPlain>
__global__ void __launch_bounds__(1024) dfs_split_kernel(const __grid_constant__ SplitParams p) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char split_smem[];
    __shared__ unsigned warp_count[32];
    __shared__ unsigned s_progress;
    const IndexRoot& ix = p.ix;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int total_warps = gridDim.x * warps_per_block;
    const int gwarp = blockIdx.x * warps_per_block + (int)warp;
    unsigned* S = reinterpret_cast<unsigned*>(split_smem + (size_t)warp * p.warp_bytes);
    short* chain = reinterpret_cast<short*>(split_smem + (size_t)warp * p.warp_bytes + p.off_chain);
    SplitState* st = p.st;

    for (int level = 0;; ++level) {
        const int m = __ldcg(&st->m), cur = __ldcg(&st->cur);
        const unsigned long long tree_count = __ldcg(&st->tree_count), log_count = __ldcg(&st->log_count);
        int reason = SEXIT_NONE;
        if (m == 0) reason = SEXIT_EMPTY;
        else if ((unsigned)m >= p.target) reason = SEXIT_TARGET;
        else if (level > 0 && !__ldcg(&st->progressed)) reason = SEXIT_STALLED;
        else if (2ull * m > __ldcg(&st->cap_pieces) || tree_count + 2ull * m > __ldcg(&st->cap_tree) ||
                 log_count + (unsigned long long)m * p.chain_cap > __ldcg(&st->cap_log)) reason = SEXIT_CAPACITY;
        else if (level >= p.max_levels) reason = SEXIT_LEVELS;
        if (reason != SEXIT_NONE) {
            if (blockIdx.x == 0 && tid == 0) st->exit_reason = reason;
            return;   // uniform: every thread decided from the same committed state
        }
        // ---- expand: one warp per piece
        for (int j = gwarp; j < m; j += total_warps) split_expand_piece<kPlain>(p, cur, j, S, chain, lane);
        grid.sync();
        // ---- ordered compaction (CTA 0): open children keep their order; dead ends forward their counts
        if (blockIdx.x == 0) {
            const SplitLevel& L = p.lv[cur];
            const SplitLevel& N = p.lv[cur ^ 1];
            unsigned* dst = p.sbuf[cur ^ 1];
            if (tid == 0) s_progress = 0u;
            __syncthreads();
            unsigned running = 0, progressed = 0;
            const int step = 4 * (int)blockDim.x;
            for (int base = 0; base < 2 * m; base += step) {
                int4 raw[4];
                unsigned cnt = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + 4 * (int)tid + u;
                    raw[u] = idx < 2 * m ? __ldcg(reinterpret_cast<const int4*>(&p.children[idx])) : make_int4(-1, 0, 0, 0);
                    cnt += raw[u].x >= 0 ? 1u : 0u;
                }
                unsigned incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned v = __shfl_up_sync(kFullMask, incl, off);
                    if ((int)lane >= off) incl += v;
                }
                if (lane == 31) warp_count[warp] = incl;
                __syncthreads();
                unsigned before = 0, total = 0;
                for (int w = 0; w < warps_per_block; ++w) {
                    const unsigned c = warp_count[w];
                    if (w < (int)warp) before += c;
                    total += c;
                }
                unsigned pos = running + before + incl - cnt;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + 4 * (int)tid + u;
                    if (idx >= 2 * m) break;
                    const int j = idx >> 1;
                    if ((u & 1) == 0) {
                        const int outcome = __ldcg(&p.parents[j].outcome);
                        if (outcome == SPLIT_DEAD) p.parents[j].fwd = pos;
                        if (outcome != SPLIT_PASS) progressed = 1u;
                    }
                    if (raw[u].x >= 0) {
                        const int outcome = __ldcg(&p.parents[j].outcome);
                        const int v = __ldcg(&p.parents[j].var);
                        BfsNode nn;
                        nn.slot = 0u; nn.n = raw[u].x; nn.kind = raw[u].y; nn.lit = raw[u].z;
                        N.meta[pos] = nn;
                        N.ptr[pos] = (raw[u].w & 1) ? nullptr : dst + (size_t)idx * ix.slot_words;
                        N.orig[pos] = __ldcg(&L.orig[j]);
                        const int32_t node = (int32_t)(tree_count + pos);
                        N.tnode[pos] = node;
                        p.t_parent[node] = __ldcg(&L.tnode[j]);
                        p.t_lit[node] = outcome == SPLIT_BRANCH ? ((u & 1) ? v : -v) : 0;
                        p.t_seg_off[node] = 0u;
                        p.t_seg_len[node] = 0u;
                        if ((u & 1) == 0) {   // the first child inherits everything charged so far + the chain
                            N.pre_nodes[pos] = __ldcg(&L.pre_nodes[j]) + __ldcg(&p.parents[j].chain_nodes);
                            N.pre_evals[pos] = __ldcg(&L.pre_evals[j]) + __ldcg(&p.parents[j].chain_evals);
                        } else {
                            N.pre_nodes[pos] = 0ull;
                            N.pre_evals[pos] = 0ull;
                        }
                        ++pos;
                    }
                }
                running += total;
                __syncthreads();
            }
            if (progressed) s_progress = 1u;
            __syncthreads();
            // dead ends: their counts go to the next piece in order if it belongs to the same
            // subproblem, else to the subproblem's tail
            for (int j = (int)tid; j < m; j += (int)blockDim.x) {
                if (__ldcg(&p.parents[j].outcome) != SPLIT_DEAD) continue;
                const unsigned fwd = p.parents[j].fwd;
                const int32_t o = __ldcg(&L.orig[j]);
                const unsigned long long dn = __ldcg(&L.pre_nodes[j]) + __ldcg(&p.parents[j].chain_nodes);
                const unsigned long long de = __ldcg(&L.pre_evals[j]) + __ldcg(&p.parents[j].chain_evals);
                if (fwd < running && N.orig[fwd] == o) {
                    atomicAdd(&N.pre_nodes[fwd], dn);
                    atomicAdd(&N.pre_evals[fwd], de);
                } else {
                    atomicAdd(&p.tail_nodes[o], dn);
                    atomicAdd(&p.tail_evals[o], de);
                }
            }
            __syncthreads();
            if (tid == 0) {
                st->m = (int)running;
                st->cur = cur ^ 1;
                st->tree_count = tree_count + running;
                st->log_count = st->log_cursor;
                st->levels_done += 1;
                st->progressed = (int)s_progress;
            }
        }
        grid.sync();
    }
}

// One thread: the choices the split made on the way to tree node `leaf`, in decision order,
// written right-aligned into out[0..cap); *out_start = index of the first one.
__global__ void split_path_kernel(const int32_t* __restrict__ t_parent, const int32_t* __restrict__ t_lit,
                                  const unsigned* __restrict__ t_seg_off, const unsigned* __restrict__ t_seg_len,
                                  const short* __restrict__ log, int32_t leaf, int32_t* __restrict__ out, int cap,
                                  int32_t* __restrict__ out_start) {
    int pos = cap;
    for (int32_t t = leaf; t >= 0; t = t_parent[t]) {
        const unsigned off = t_seg_off[t], len = t_seg_len[t];
        for (int i = (int)len - 1; i >= 0; --i) {
            --pos;
            if (pos >= 0) out[pos] = (int32_t)log[off + i];
        }
        if (t_lit[t] != 0) {
            --pos;
            if (pos >= 0) out[pos] = t_lit[t];
        }
    }
    *out_start = pos;
}

}  // namespace ndp
