// ndp_index.cuh -- the occurrence-index engine (BFS levels and DFS).
//
// Same search as the scan engine (ndp_kernels.cuh) -- the reference's Satisfy_iterative_BFS /
// Satisfy_iterative node for node (NDP-5_6_7.cpp:881-944, 784-878), same choices in the same
// order, same node and clause-evaluation counts -- but a node no longer scans its clause set to
// find the ~dozen clauses that mention the assigned variable.  Every node of a run is the ROOT
// clause array filtered by an assignment (SURVEY.md §8a equivalences), so one read-only index
// over the root serves them all:
//
//   root[c]      the packed root clause (ndp_clause.cuh record; zc = zeros it was parsed with)
//   occ[v]       clauses that mention variable v, merged per clause: c | n_pos << 24 | n_neg << 26
//                (n_pos / n_neg = how many of the clause's literals are +v / -v, so duplicate
//                literals and tautologies need no special case)
//
// and a node's whole state ("state slot") is two bit-planes over root clause ids plus the assignment:
//
//   P0, P1            2 bits per clause, (P1,P0): 00 gone (satisfied), 01 live with no zero literal,
//                     10 live with one, 11 live with two.  In the text below B[k] = "live with exactly k
//                     zeros": B[2] = P1 & P0 (the unit clauses), B[1] = P1 & ~P0, B[0] = ~P1 & P0
//   table             2 bits per variable (0 unassigned, 1 true, 2 false)
//   S2                one bit per word of B[2] ("this word is non-zero"): the next unit clause -- what
//                     99.6 % of the reference's nodes ask for -- is two find-first-sets away
//
// choice() (NDP:753-781) becomes find-first-set: the first live two-zero clause is the lowest set
// bit of B[2], the first one-zero clause that of B[1], A[0] that of B[0].  Resolution
// (NDP:654-697 / 700-750) becomes: walk occ[v]; a clause holding the true literal leaves its
// bitmap (satisfied, dropped); one holding only false literals moves from B[k] to B[k + n_false],
// and reaching three zeros is the conflict.  The materialised clause list of a node (needed only
// for -s JSON, ndp_frontier_get and the scan engine) is the root filtered by its table.
#pragma once
#include <cooperative_groups.h>

#include "ndp_clause.cuh"
#include "ndp_kernels.cuh"

namespace ndp {

struct IndexRoot {
    const pclause* root;              // [n_root]
    int n_root;
    // Occurrence rows, one per variable, MERGED PER BITMAP WORD: entry = {w, pos, neg, 0} says "of the 32
    // clauses of plane word w, those in `pos` hold +v once, those in `neg` hold -v once".  A resolution
    // step is then plain word arithmetic on (P0[w], P1[w]) by ONE lane per word -- no per-clause atomics,
    // no two lanes on one word.  row_hdr[v] = {first entry, entry count | multi << 31}.
    const uint2* row_hdr;             // [max_var + 1]
    const uint4* rows;
    // Clauses that mention a variable more than once ((x, x, y), (x, -x, y): duplicates and tautologies)
    // are left out of that variable's row and listed here per clause: c | n_pos << 24 | n_neg << 26
    // (n_pos / n_neg = how many of the clause's literals are +v / -v).  Empty for every generated CNF.
    const unsigned* occ_off;          // [max_var + 2]
    const unsigned* occ;
    int bm_words, table_words;
    int table_off;                    // = 2 * bm_words: the slot is [P0 | P1 | table | S2]
    int sum_words, sum_off;           // summary of B[2]: bit w of word sum_off + (w >> 5) <=> B[2][w] != 0  (sum_words <= 32, or 0: none)
    int slot_words;                   // 2 * bm_words + table_words + sum_words, rounded up to a multiple of 4
    int n_zero3;                      // root clauses that are {0,0,0} (never leave, always conflict)
    int first_zero3;                  // lowest such index, or -1
    int plain;                        // sum_words > 0, n_zero3 == 0 and occ is empty: kernels may use their kPlain instantiation
};

// lowest clause of class k (k = 1: one zero literal, k = 0: none) in the two planes, or -1
__device__ __forceinline__ int planes_first(const IndexRoot& ix, const unsigned* S, int k, unsigned lane) {
    const unsigned* P0 = S;
    const unsigned* P1 = S + ix.bm_words;
    for (int w0 = 0; w0 < ix.bm_words; w0 += 32) {
        const int w = w0 + (int)lane;
        unsigned word = 0u;
        if (w < ix.bm_words) { const unsigned a = P0[w], b = P1[w]; word = k == 1 ? (b & ~a) : k == 0 ? (~b & a) : (b & a); }
        const unsigned m = __ballot_sync(kFullMask, word != 0u);
        if (m) {
            const int src = __ffs(m) - 1;
            const unsigned wd = __shfl_sync(kFullMask, word, src);
            return (w0 + src) * 32 + __ffs(wd) - 1;
        }
    }
    return -1;
}

// Derive the two planes (and the summary) from the assignment table with one pass over the root.
// Returns the number of live clauses (all-zero root clauses included: they are part of |A|).
__device__ __noinline__ int index_rebuild(const IndexRoot& ix, unsigned* B, const unsigned* table, unsigned lane) {
    int n_live = 0;
    unsigned sum_acc = 0;   // summary bits of the current group of 32 bitmap words
    for (int base = 0; base < ix.bm_words * 32; base += 32) {
        const int c = base + (int)lane;
        bool live = false;
        unsigned zc = 0;
        if (c < ix.n_root) {
            const pclause cl = __ldg(&ix.root[c]);
            const unsigned f[3] = {cl.x & 0xffffu, cl.x >> 16, cl.y & 0xffffu};
            bool sat = false;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (f[j] == 0u) ++zc;
                else {
                    const unsigned val = table_get(table, f[j] >> 1);
                    const unsigned truth = (f[j] & 1u) ? 2u : 1u;
                    if (val == truth) sat = true;
                    else if (val != 0u) ++zc;
                }
            }
            live = !sat;
        }
        const unsigned b0 = __ballot_sync(kFullMask, live && zc == 0u);
        const unsigned b1 = __ballot_sync(kFullMask, live && zc == 1u);
        const unsigned b2 = __ballot_sync(kFullMask, live && zc == 2u);
        const unsigned b3 = __ballot_sync(kFullMask, live && zc == 3u);
        const int w = base >> 5;
        if (lane == 0) {
            B[w] = b0 | b2;                     // P0
            B[ix.bm_words + w] = b1 | b2;       // P1
        }
        if (b2) sum_acc |= 1u << (w & 31);
        if (ix.sum_words && lane == 0 && ((w & 31) == 31 || w == ix.bm_words - 1)) { B[ix.sum_off + (w >> 5)] = sum_acc; }
        if ((w & 31) == 31) sum_acc = 0u;
        n_live += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
    }
    __syncwarp();
    return n_live;
}

// The literal choice() takes from live clause c: the LAST of its literals that is still there
// (non-zero and unassigned).  Lanes 0..2 test one literal each.
__device__ __forceinline__ int clause_last_literal(const IndexRoot& ix, const unsigned* table, int c, unsigned lane) {
    const pclause cl = __ldg(&ix.root[c]);
    unsigned f = ((lane >= 2u ? cl.y : cl.x) >> ((lane & 1u) << 4)) & 0xffffu;   // code of literal `lane`
    if (lane > 2u) f = 0u;
    // f = var << 1 | sign: table word var >> 4 = f >> 5, bit pair (var & 15) * 2 = f & 30
    const bool there = f != 0u && ((table[f >> 5] >> (f & 30u)) & 3u) == 0u;
    const unsigned m = __ballot_sync(kFullMask, there);
    return code_lit(__shfl_sync(kFullMask, f, 31 - __clz((int)m)));
}

// choice() when no clause has two zeros (NDP:767-779): a decision.  Rare (0.4 % of the DFS nodes).
__device__ __noinline__ int index_choice_decide(const IndexRoot& ix, const unsigned* B, const unsigned* table, unsigned lane) {
    int c = planes_first(ix, B, 1, lane);
    if (c >= 0) {
        const int l = clause_last_literal(ix, table, c, lane);
        return l < 0 ? -l : l;
    }
    // neither units nor one-zero clauses: |A[0].l[0]| (NDP:778-779)
    c = planes_first(ix, B, 0, lane);
    if (ix.first_zero3 >= 0 && (c < 0 || ix.first_zero3 < c)) return 0;   // A[0] is {0,0,0}
    const int l = code_lit(__ldg(&ix.root[c]).x & 0xffffu);
    return l < 0 ? -l : l;
}

// What choice() returns for the current state: kind/lit as in SetInfo.
// kPlain (here and below): the root has a B[2] summary, no {0,0,0} clause and no clause that mentions a
// variable twice -- true for every generated factoring CNF -- so the tests for those cases are compiled out.
template <bool kPlain = false>
__device__ __forceinline__ void index_choice(const IndexRoot& ix, const unsigned* B, const unsigned* table, int n_live,
                                             unsigned lane, int& kind, int& lit) {
    if (!kPlain && n_live == 0) { kind = K_EMPTY; lit = 0; return; }
    int c;
    if (kPlain || ix.sum_words) {
        // first unit clause through the summary: one ballot over <= 32 summary words, then one word of B[2]
        const unsigned sw = (int)lane < ix.sum_words ? B[ix.sum_off + lane] : 0u;
        const unsigned m = __ballot_sync(kFullMask, sw != 0u);
        if (m) {
            const int src = __ffs(m) - 1;
            const int w = src * 32 + __ffs(__shfl_sync(kFullMask, sw, src)) - 1;
            c = w * 32 + __ffs(B[w] & B[ix.bm_words + w]) - 1;
        } else c = -1;
    } else c = planes_first(ix, B, 2, lane);
    if (c >= 0) { kind = K_UNIT; lit = clause_last_literal(ix, table, c, lane); }
    else { kind = K_DECIDE; lit = index_choice_decide(ix, B, table, lane); }
}

// The per-clause pass over the clauses that mention |t| more than once (IndexRoot::occ).  Per-lane partial
// results: `dropped` clauses satisfied, `confl` a clause reached three zeros.  kWrite applies the step.
template <bool kWrite>
__device__ __noinline__ void index_assign_multi(const IndexRoot& ix, unsigned* B, int t, unsigned lane, unsigned& dropped,
                                                bool& confl) {
    const unsigned v = (unsigned)(t < 0 ? -t : t);
    const unsigned a = __ldg(&ix.occ_off[v]), b = __ldg(&ix.occ_off[v + 1]);
    for (unsigned base = a; base < b; base += 32) {   // warp-uniform trip count (the summary fix-up has a warp barrier)
        const unsigned i = base + lane;
        const unsigned e = i < b ? __ldg(&ix.occ[i]) : 0u;
        const unsigned c = e & 0xffffffu, n_pos = (e >> 24) & 3u, n_neg = (e >> 26) & 3u;
        const unsigned w = c >> 5, bit = 1u << (c & 31u);
        int k = -1;
        unsigned p0 = 0u, p1 = 0u;
        if (i < b) {
            p0 = B[w] & bit;
            p1 = B[ix.bm_words + w] & bit;
            k = (p1 ? 2 : 0) + (p0 ? 1 : 0) - 1;   // -1: already satisfied earlier on this path
        }
        bool touch2 = false;   // this lane changed word w of B[2] = P1 & P0
        if (k >= 0) {
            const unsigned n_true = t > 0 ? n_pos : n_neg, n_false = t > 0 ? n_neg : n_pos;
            const int k2 = n_true ? -1 : k + (int)n_false;      // class afterwards; -1 / >= 3: the clause leaves the planes
            if (n_true) ++dropped;
            else if (k2 >= 3) confl = true;
            if (kWrite && k2 != k) {
                const unsigned code2 = (k2 >= 0 && k2 < 3) ? (unsigned)(k2 + 1) : 0u;
                if ((p0 != 0u) != ((code2 & 1u) != 0u)) atomicXor(&B[w], bit);                  // only this lane owns the bit
                if ((p1 != 0u) != ((code2 & 2u) != 0u)) atomicXor(&B[ix.bm_words + w], bit);
                touch2 = k == 2 || code2 == 3u;
            }
        }
        if (kWrite && ix.sum_words) {
            // bring the summary of B[2] in line with the words this step changed: every lane that
            // touched a word reads its FINAL value (all bitmap atomics of the step are done after the
            // warp barrier), so lanes sharing a word agree on what to write
            __syncwarp();
            if (touch2) {
                unsigned* sp = &B[ix.sum_off + (w >> 5)];
                const unsigned sbit = 1u << (w & 31u);
                if (B[w] & B[ix.bm_words + w]) atomicOr(sp, sbit);
                else atomicAnd(sp, ~sbit);
            }
        }
    }
    if (kWrite) __syncwarp();
}

constexpr unsigned kConflBit = 1u << 26;   // per-lane conflict flag riding on the dropped count (n_root < 2^24) through one redux

// One row entry of a resolution step: e = {w, pos, neg, -}.  Returns the entry's contribution to the
// step's result: clauses satisfied (+ kConflBit if one of the word's clauses reached three zeros).
template <bool kWrite, bool kPlain>
__device__ __forceinline__ unsigned index_assign_entry(const IndexRoot& ix, unsigned* B, const uint4 e, bool positive) {
    const unsigned T = positive ? e.y : e.z, F = positive ? e.z : e.y;
    unsigned* w0 = B + e.x;
    unsigned* w1 = w0 + ix.bm_words;
    unsigned p0 = *w0, p1 = *w1;
    const unsigned live = p0 | p1, old2 = p0 & p1;
    const unsigned f = F & live;
    unsigned acc = (unsigned)__popc(T & live);
    if (f & old2) acc |= kConflBit;
    if (kWrite) {
        p0 &= ~T;
        p1 &= ~T;
        const unsigned carry = p0 & f;   // (P1,P0) += 1 on the F bits: 01 -> 10 -> 11 -> (00 + conflict)
        p0 ^= f;
        p1 ^= carry;
        *w0 = p0;
        *w1 = p1;
        if ((kPlain || ix.sum_words) && ((old2 != 0u) != ((p0 & p1) != 0u))) {
            // the word gained its first / lost its last unit clause: one summary bit, owned by this lane
            unsigned* sp = &B[ix.sum_off + (e.x >> 5)];
            const unsigned sbit = 1u << (e.x & 31u);
            if (old2 == 0u) atomicOr(sp, sbit);
            else atomicAnd(sp, ~sbit);
        }
    }
    return acc;
}

// Resolution for "literal t becomes true" (NDP:654-697 / 700-750 on the planes).  kWrite applies it;
// otherwise it only classifies.  Returns the live count afterwards; conflict is raised when a clause
// reaches three zeros.  One lane per touched plane word: satisfied clauses (T) leave both planes, clauses
// holding the false literal (F) move up one class -- a 2-bit increment (P1,P0) += 1 on the F bits, whose
// overflow out of 11 is the conflict.
template <bool kWrite, bool kPlain = false>
__device__ __forceinline__ int index_assign(const IndexRoot& ix, unsigned* B, int n_live, int t, bool& conflict,
                                            unsigned lane) {
    const unsigned v = (unsigned)(t < 0 ? -t : t);
    const uint2 hdr = __ldg(&ix.row_hdr[v]);
    const unsigned cnt = hdr.y & 0x7fffffffu;
    unsigned acc = 0;   // satisfied clauses + kConflBit per conflict
    if (lane < cnt) acc = index_assign_entry<kWrite, kPlain>(ix, B, __ldg(&ix.rows[hdr.x + lane]), t > 0);
    if (cnt > 32u) {    // rows longer than a warp: the multiplier's input bits (decisions only)
#pragma unroll 1
        for (unsigned i = lane + 32u; i < cnt; i += 32u) {
            const unsigned a2 = index_assign_entry<kWrite, kPlain>(ix, B, __ldg(&ix.rows[hdr.x + i]), t > 0);
            acc = (acc + (a2 & (kConflBit - 1u))) | (a2 & kConflBit);
        }
    }
    if (kWrite) __syncwarp();
    if (!kPlain && (hdr.y >> 31)) {   // the variable also sits in duplicate-literal / tautological clauses: per-clause pass
        unsigned dropped = 0;
        bool confl = false;
        index_assign_multi<kWrite>(ix, B, t, lane, dropped, confl);
        acc += dropped;
        if (confl) acc |= kConflBit;
    }
    const unsigned total = __reduce_add_sync(kFullMask, acc);   // <= 32 flags of 2^26 stay below 2^32
    conflict = (!kPlain && ix.n_zero3 > 0) || (total >> 26) != 0u;
    return n_live - (int)(total & (kConflBit - 1u));
}

// Both children of a decision on variable `var` in one read-only pass: LA = var true, RA = var false.
__device__ __forceinline__ void index_classify(const IndexRoot& ix, unsigned* B, int n_live, int var, int& la_n, bool& la_conf,
                                               int& ra_n, bool& ra_conf, unsigned lane) {
    const uint2 hdr = __ldg(&ix.row_hdr[var]);
    const unsigned cnt = hdr.y & 0x7fffffffu;
    unsigned la = 0, ra = 0;
    for (unsigned base = 0; base < cnt; base += 32) {
        const unsigned i = base + lane;
        if (i < cnt) {
            const uint4 e = __ldg(&ix.rows[hdr.x + i]);
            const unsigned p0 = B[e.x], p1 = B[ix.bm_words + e.x];
            const unsigned live = p0 | p1, two = p0 & p1;
            la += (unsigned)__popc(e.y & live);
            ra += (unsigned)__popc(e.z & live);
            if (e.z & two) la |= kConflBit;
            if (e.y & two) ra |= kConflBit;
        }
    }
    if (hdr.y >> 31) {
        unsigned d = 0;
        bool c = false;
        index_assign_multi<false>(ix, B, var, lane, d, c);
        la += d;
        if (c) la |= kConflBit;
        d = 0; c = false;
        index_assign_multi<false>(ix, B, -var, lane, d, c);
        ra += d;
        if (c) ra |= kConflBit;
    }
    const unsigned tl = __reduce_add_sync(kFullMask, la), tr = __reduce_add_sync(kFullMask, ra);
    la_conf = ix.n_zero3 > 0 || (tl >> 26) != 0u;
    ra_conf = ix.n_zero3 > 0 || (tr >> 26) != 0u;
    la_n = n_live - (int)(tl & (kConflBit - 1u));
    ra_n = n_live - (int)(tr & (kConflBit - 1u));
}

__device__ __forceinline__ void table_set(unsigned* table, int t) {
    const unsigned v = (unsigned)(t < 0 ? -t : t);
    table[v >> 4] |= (t > 0 ? 1u : 2u) << ((v & 15u) * 2u);
}

// copy one state slot (slot_words is a multiple of 4; both sides 16-byte aligned)
__device__ __forceinline__ void slot_copy(unsigned* dst, const unsigned* src, int slot_words, unsigned lane) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = lane; i < slot_words / 4; i += 32) d[i] = s[i];
}
// same, reading the source through L2 only (.cg): used where another SM may have written the slot
// earlier in the same launch
__device__ __forceinline__ void slot_load_cg(unsigned* dst_shared, const unsigned* src, int slot_words, unsigned lane) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst_shared);
    for (int i = lane; i < slot_words / 4; i += 32) d[i] = __ldcg(&s[i]);
}

// ------------------------------------------------------------------------------------------
// BFS level over state slots: one warp per node (Satisfy_iterative_BFS loop body, NDP:898-925).
// The parent's slot is staged in shared memory, the branch is applied there, and an open child
// (non-empty, conflict-free: NDP:908,919) is written to slot 2j (LA) / 2j+1 (RA) of the other
// buffer together with what choice() returns for it.  Per-child metadata as in bfs_expand_kernel.
// One node: stage the parent's slot in S (shared), apply each attempted branch, write open
// children to dst slots 2j / 2j+1.  ch[0] = LA, ch[1] = RA.
template <bool kPlain = false>
__device__ __forceinline__ void index_expand_node(const IndexRoot& ix, const unsigned* src,   // no __restrict__:
                                                  unsigned* dst, int j, const BfsNode nd, unsigned* S,   // a persistent kernel re-reads what it wrote

                                                  unsigned lane, BfsChild ch[2]) {
    unsigned* table = S + ix.table_off;
    const unsigned* parent = src + (size_t)nd.slot * ix.slot_words;
    ch[0].n = ch[1].n = -1;
    ch[0].kind = ch[1].kind = K_EMPTY;
    ch[0].lit = ch[1].lit = 0;
    ch[0].empty = ch[1].empty = 0;
    const int var = nd.kind == K_UNIT ? (nd.lit < 0 ? -nd.lit : nd.lit) : nd.lit;
    if (nd.n <= 0 || var == 0) return;
    for (int b = 0; b < 2; ++b) {
        const int t = b == 0 ? var : -var;   // b=0: LA (i := true), b=1: RA (i := false)
        if (nd.kind == K_UNIT && t != nd.lit) continue;   // falsifies the unit clause: {0,0,0}
        __syncwarp();
        slot_load_cg(S, parent, ix.slot_words, lane);
        __syncwarp();
        if (lane == 0) table_set(table, t);
        bool conflict;
        const int n_after = index_assign<true, kPlain>(ix, S, nd.n, t, conflict, lane);
        if (n_after == 0) ch[b].empty = 1;
        if (n_after > 0 && !conflict) {
            int kind, lit;
            index_choice<kPlain>(ix, S, table, n_after, lane, kind, lit);
            ch[b].n = n_after; ch[b].kind = kind; ch[b].lit = lit;
            slot_copy(dst + ((size_t)2 * j + b) * ix.slot_words, S, ix.slot_words, lane);
        }
    }
}

__global__ void __launch_bounds__(256) bfs_index_expand_kernel(const IndexRoot ix, const unsigned* __restrict__ src,
                                                               unsigned* __restrict__ dst,
                                                               const BfsNode* __restrict__ nodes, int m,
                                                               BfsChild* __restrict__ children) {
    extern __shared__ __align__(16) unsigned smem_words[];
    const unsigned lane = threadIdx.x & 31u;
    const int warps_per_block = blockDim.x >> 5;
    unsigned* S = smem_words + (size_t)(threadIdx.x >> 5) * ix.slot_words;
    for (int j = blockIdx.x * warps_per_block + (threadIdx.x >> 5); j < m; j += gridDim.x * warps_per_block) {
        BfsChild ch[2];
        index_expand_node(ix, src, dst, j, nodes[j], S, lane, ch);
        if (lane == 0) { children[2 * j] = ch[0]; children[2 * j + 1] = ch[1]; }
    }
}

// The whole level loop on the device (cooperative launch, one CTA per SM), ONE grid.sync per step.
// A step advances the queue by K levels (a WINDOW) while no stop test can fire, by one level with
// the exact tests otherwise.  CTA b owns the contiguous node range [b*B, (b+1)*B) of the level:
//   expand    one warp per node.  K = 1: index_expand_node -> two children.  K > 1:
//             index_window_node walks the node's subtree K levels down, LA before RA -- the order
//             in which those descendants sit in the FIFO queue K levels later (NDP:904-925) -- and
//             emits up to kWindowFan of them with their K choices; per-level expansion / push /
//             clause-evaluation counts are accumulated for the reference's counters
//   scan      block-wide exclusive scan of the per-node output counts: local queue positions
//   publish   the CTA's totals, tagged with the step
//   look-back warp 0 waits for the totals of CTAs 0..b-1 (all co-resident) -> the CTA's base position
//   emit      next node list in queue order, choice-tree entries; K = 1 also evaluates the exact -q
//             test "queue size >= Q before this pop" (NDP:894) at EVERY node
//   grid.sync
//   commit    every CTA derives the same next state from the same published totals.  A window is
//             committed only if none of its K levels could have fired a stop test (-q: 2 m < Q at
//             every level; -d: iterations + m < D) and no node overflowed its fan-out; otherwise it
//             is redone, shorter, from the untouched source level.
// The kernel returns to the host only when it has to: the level in which a stop test fires (left
// UNTOUCHED for the host's exact walk), an empty queue, or arenas that must grow.
enum : int { PEXIT_NONE = 0, PEXIT_EMPTY = 1, PEXIT_HOST_LEVEL = 2, PEXIT_CAPACITY = 3, PEXIT_LEVELS = 4,
              PEXIT_WIDE = 5 };   // WIDE: the level has more nodes per CTA than the shared scan arrays hold
constexpr int kPersistMaxBlockNodes = 2048;   // nodes of one level a CTA can own (shared-memory scan arrays)
constexpr int kWindowMax = 32;                // levels per window
constexpr int kWindowFan = 8;                 // descendants one node may emit per window of > 16 levels (4 up to 16 levels); else the window is redone
constexpr int kWindowPend = 4;                // pending RA siblings one warp can hold (snapshots in global scratch)
struct PersistState {
    int m, cur, cur_buf, levels_done, exit_reason, window;
    long long it, task_count;
    unsigned long long evals;
    unsigned long long tree_count;
    // capacities (set by the host)
    unsigned long long cap_slots, cap_nodes, cap_tree;
    unsigned long long windows_ok, windows_redone;
    unsigned long long t_phase[5];   // CTA 0, ns: expand | scan + publish + look-back | emit | grid.sync | commit
};
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct PersistTotals {            // one per (step parity, CTA): what the look-back needs; `tag` is written last
    unsigned children, tag;
};
struct PersistAcc {               // grid-wide sums of one step, accumulated with atomics (three in rotation)
    unsigned children, overflow;
    unsigned expanded[kWindowMax], pushed[kWindowMax], branched[kWindowMax];
    unsigned long long evals[kWindowMax];
};
struct PersistScratch {           // zeroed by the host before every launch
    PersistTotals totals[2][160];
    PersistAcc acc[3];
    unsigned stop_tag[2];
};

// One node, K levels down (K > 1).  Outputs e = 0..count-1 in queue order: state slot
// dst[kWindowFan * j + e], metadata children[kWindowFan * j + e], choices paths[(kWindowFan * j + e) * kWindowMax + r].
// Returns the output count, or -1 if the node needs more than kWindowFan outputs / kWindowPend snapshots.
template <bool kPlain = false>
__device__ __forceinline__ int index_window_node(const IndexRoot& ix, const unsigned* src, unsigned* dst, int j,
                                                 const BfsNode nd, unsigned* S, unsigned* snap, unsigned lane, int K, int fan,
                                                 BfsChild* children, short* paths, unsigned& my_exp, unsigned& my_push,
                                                 unsigned& my_br, unsigned long long& my_ev) {
    // my_*: lane r accumulates the counters of relative level r (summed over the warp's nodes by the caller)
    unsigned* table = S + ix.table_off;
    __syncwarp();
    slot_load_cg(S, src + (size_t)nd.slot * ix.slot_words, ix.slot_words, lane);
    __syncwarp();
    int n_live = nd.n, kind = nd.kind, lit = nd.lit;
    int r = 0, out = 0, npend = 0;
    int my_path = 0;                      // lane r holds the choice made at relative level r
    int pend_r = 0, pend_var = 0, pend_n = 0;   // lane p holds pending entry p
    for (;;) {
        bool backtrack = false;
        if (r == K) {
            if (out == fan) return -1;
            const size_t o = (size_t)fan * j + out;
            slot_copy(dst + o * ix.slot_words, S, ix.slot_words, lane);
            if ((int)lane < K) paths[o * kWindowMax + lane] = (short)my_path;
            if (lane == 0) { BfsChild c; c.n = n_live; c.kind = kind; c.lit = lit; c.empty = 0; children[o] = c; }
            ++out;
            backtrack = true;
        } else {
            const int var = kind == K_UNIT ? (lit < 0 ? -lit : lit) : lit;
            if (n_live <= 0 || var == 0) {
                backtrack = true;         // choice() == 0: popped, nothing pushed, no iteration (NDP:901-902)
            } else {
                if ((int)lane == r) { ++my_exp; my_ev += (unsigned long long)n_live; }
                int t = 0;
                if (kind == K_UNIT) {
                    t = lit;              // the other branch falsifies the unit clause: {0,0,0}, never pushed
                } else {
                    bool la_conf, ra_conf;
                    int la_n, ra_n;
                    index_classify(ix, S, n_live, var, la_n, la_conf, ra_n, ra_conf, lane);
                    const bool la_open = la_n > 0 && !la_conf, ra_open = ra_n > 0 && !ra_conf;   // NDP:908,919
                    if (la_open && ra_open) {
                        if (npend == kWindowPend) return -1;
                        __syncwarp();
                        slot_copy(snap + (size_t)npend * ix.slot_words, S, ix.slot_words, lane);
                        if ((int)lane == npend) { pend_r = r; pend_var = var; pend_n = n_live; }
                        ++npend;
                        if ((int)lane == r) { ++my_push; ++my_br; }   // RA is pushed now, visited later
                        t = var;
                    } else if (la_open) t = var;
                    else if (ra_open) t = -var;
                }
                if (t == 0) {
                    backtrack = true;
                } else {
                    __syncwarp();
                    if (lane == 0) table_set(table, t);
                    bool conflict;
                    const int n2 = index_assign<true, kPlain>(ix, S, n_live, t, conflict, lane);
                    if (n2 > 0 && !conflict) {
                        if ((int)lane == r) ++my_push;   // NDP:912,923
                        if ((int)lane == r) my_path = t;
                        n_live = n2;
                        index_choice<kPlain>(ix, S, table, n_live, lane, kind, lit);
                        ++r;
                    } else {
                        backtrack = true; // empty (a dropped solution leaf) or conflict: not pushed
                    }
                }
            }
        }
        if (!backtrack) continue;
        if (npend == 0) break;
        // the deepest pending RA sibling: restore the state before its branch, take i := false
        --npend;
        const int pr = __shfl_sync(kFullMask, pend_r, npend);
        const int pv = __shfl_sync(kFullMask, pend_var, npend);
        const int pn = __shfl_sync(kFullMask, pend_n, npend);
        __syncwarp();
        slot_load_cg(S, snap + (size_t)npend * ix.slot_words, ix.slot_words, lane);
        __syncwarp();
        if (lane == 0) table_set(table, -pv);
        bool conflict;
        n_live = index_assign<true, kPlain>(ix, S, pn, -pv, conflict, lane);   // known open and non-empty
        if ((int)lane == pr) my_path = -pv;
        index_choice<kPlain>(ix, S, table, n_live, lane, kind, lit);
        r = pr + 1;
    }
    return out;
}

template <bool kPlain>
__global__ void __launch_bounds__(1024) bfs_index_persistent_kernel(const IndexRoot ix, unsigned* sbuf0, unsigned* sbuf1,
                                                                   BfsNode* nodes0, BfsNode* nodes1, int32_t* ntree0,
                                                                   int32_t* ntree1, BfsChild* children,
                                                                   int32_t* tree_parent, int32_t* tree_lit, int D, int Q,
                                                                   int ovr, int max_levels, PersistState* st,
                                                                   PersistScratch* scratch, unsigned* snaps, short* paths,
                                                                   int window_max, int max_block_nodes) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned smem_words[];
    __shared__ unsigned char s_cnt[kPersistMaxBlockNodes];        // K = 1: bit 0 LA open, bit 1 RA open; K > 1: outputs
    __shared__ unsigned short s_off[kPersistMaxBlockNodes];       // outputs of the CTA's earlier nodes
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_exp[kWindowMax], s_push[kWindowMax], s_br[kWindowMax];
    __shared__ unsigned long long s_ev[kWindowMax];
    __shared__ unsigned s_children, s_overflow, s_base, s_stop;
    __shared__ unsigned t_exp[kWindowMax], t_push[kWindowMax], t_br[kWindowMax], t_children, t_overflow;
    __shared__ unsigned long long t_ev[kWindowMax];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    unsigned* S = smem_words + (size_t)warp * ix.slot_words;
    unsigned* snap = snaps ? snaps + ((size_t)b * warps_per_block + warp) * kWindowPend * ix.slot_words : nullptr;

    // the committed state, carried in registers by every thread (identical across the grid)
    int m = st->m, cur = st->cur, cur_buf = st->cur_buf, levels_done = st->levels_done;
    long long it = st->it, task_count = st->task_count;
    unsigned long long evals_total = st->evals, tree_count = st->tree_count;
    const unsigned long long cap_slots = st->cap_slots, cap_nodes = st->cap_nodes, cap_tree = st->cap_tree;
    unsigned long long windows_ok = 0, windows_redone = 0;
    unsigned long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0, tmark = 0;   // phase clocks (thread 0 of CTA 0)
    const bool clocked = b == 0 && tid == 0;
    int reason = PEXIT_NONE;
    const int Kmax = (snaps && paths) ? min(max(window_max, 1), kWindowMax) : 1;
    int Kcur = Kmax;
    bool near_stop = false;   // a window was cut short by the stop rules: finish level by level

    for (int step = 0;; ++step) {
        if (m == 0) reason = PEXIT_EMPTY;
        else if ((Q > 0 && m >= Q) || (D == -1 && !ovr)) reason = PEXIT_HOST_LEVEL;          // NDP:894-897 at the level's first pop
        else if (Q <= 0 && it + m >= (long long)D) reason = PEXIT_HOST_LEVEL;                // the -d budget ends inside this level
        else if ((m + G - 1) / G > max_block_nodes) reason = PEXIT_WIDE;   // this level goes through the per-level kernels
        else if (2ull * m > cap_slots || 2ull * m > cap_nodes || tree_count + 2ull * m > cap_tree) reason = PEXIT_CAPACITY;
        else if (levels_done >= max_levels) reason = PEXIT_LEVELS;
        if (reason != PEXIT_NONE) break;   // uniform: every thread holds the same state
        // window length of this step: K levels need kWindowFan slots / node entries per node and K tree entries per output
        int K = Kcur;
        auto fits = [&](int k) {   // k levels need fan(k) slots / node entries per node and k tree entries per output
            const unsigned long long fk = k > 16 ? kWindowFan : 4;
            return fk * m <= cap_slots && fk * m <= cap_nodes && tree_count + fk * m * k <= cap_tree;
        };
        if (K > 16 && !fits(K)) K = 16;
        if (K > 1 && !fits(K)) K = 1;
        const int fan = K > 16 ? kWindowFan : K > 1 ? 4 : 2;
        const unsigned tag = (unsigned)step + 1u;
        const int par = step & 1;
        unsigned* src = cur_buf ? sbuf1 : sbuf0;
        unsigned* dst = cur_buf ? sbuf0 : sbuf1;
        const BfsNode* nodes = cur ? nodes1 : nodes0;
        const int32_t* ntree = cur ? ntree1 : ntree0;
        BfsNode* nnodes = cur ? nodes0 : nodes1;
        int32_t* nntree = cur ? ntree0 : ntree1;
        const int B = (m + G - 1) / G;
        const int j0 = b * B;
        const int nb = max(0, min(B, m - j0));         // nodes this CTA owns
        if (tid < kWindowMax) { s_exp[tid] = 0u; s_push[tid] = 0u; s_br[tid] = 0u; s_ev[tid] = 0ull; }
        if (tid == 0) { s_children = 0u; s_overflow = 0u; }
        __syncthreads();
        if (clocked) tmark = global_ns();
        // ---- expand
        unsigned my_exp = 0, my_push = 0, my_br = 0;   // window mode: lane r counts relative level r
        unsigned long long my_ev = 0;
        for (int k = (int)warp; k < nb; k += warps_per_block) {
            const int j = j0 + k;
            BfsNode nd;
            {
                const int4 raw = __ldcg(reinterpret_cast<const int4*>(&nodes[j]));
                nd.slot = (uint32_t)raw.x; nd.n = raw.y; nd.kind = raw.z; nd.lit = raw.w;
            }
            if (K == 1) {
                BfsChild ch[2];
                index_expand_node<kPlain>(ix, src, dst, j, nd, S, lane, ch);
                if (lane == 0) {
                    children[2 * j] = ch[0];
                    children[2 * j + 1] = ch[1];
                    s_cnt[k] = (unsigned char)((ch[0].n >= 0 ? 1 : 0) | (ch[1].n >= 0 ? 2 : 0));
                    const int var = nd.kind == K_UNIT ? (nd.lit < 0 ? -nd.lit : nd.lit) : nd.lit;
                    if (nd.n > 0 && var != 0) { atomicAdd(&s_exp[0], 1u); atomicAdd(&s_ev[0], (unsigned long long)nd.n); }
                }
            } else {
                const int c = index_window_node<kPlain>(ix, src, dst, j, nd, S, snap, lane, K, fan, children, paths, my_exp, my_push, my_br, my_ev);
                if (lane == 0) {
                    s_cnt[k] = (unsigned char)(c < 0 ? 0 : c);
                    if (c < 0) s_overflow = 1u;
                }
            }
        }
        if (K > 1 && (int)lane < K && my_exp) {
            atomicAdd(&s_exp[lane], my_exp);
            atomicAdd(&s_push[lane], my_push);
            atomicAdd(&s_br[lane], my_br);
            atomicAdd(&s_ev[lane], my_ev);
        }
        __syncthreads();
        if (clocked) { const unsigned long long t = global_ns(); tp0 += t - tmark; tmark = t; }
        // ---- block-wide exclusive scan of the per-node output counts (contiguous chunk per thread)
        {
            const int chunk = (nb + (int)blockDim.x - 1) / (int)blockDim.x;
            const int k0 = (int)tid * chunk, k1 = min(nb, k0 + chunk);
            unsigned sum = 0;
            for (int k = k0; k < k1; ++k) sum += K == 1 ? (unsigned)__popc((unsigned)s_cnt[k]) : (unsigned)s_cnt[k];
            unsigned incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned v = __shfl_up_sync(kFullMask, incl, off);
                if ((int)lane >= off) incl += v;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            unsigned before = 0;
            for (int w = 0; w < (int)warp; ++w) before += s_warp[w];
            unsigned run = before + incl - sum;
            for (int k = k0; k < k1; ++k) {
                s_off[k] = (unsigned short)run;
                run += K == 1 ? (unsigned)__popc((unsigned)s_cnt[k]) : (unsigned)s_cnt[k];
            }
            if (tid == blockDim.x - 1) s_children = run;
        }
        __syncthreads();
        // ---- publish this CTA's totals, then look back over the earlier CTAs
        if (warp == 0) {
            PersistTotals* mine = &scratch->totals[par][b];
            PersistAcc* acc = &scratch->acc[step % 3];
            if ((int)lane < K && s_exp[lane]) {
                atomicAdd(&acc->expanded[lane], s_exp[lane]);
                if (s_push[lane]) atomicAdd(&acc->pushed[lane], s_push[lane]);
                if (s_br[lane]) atomicAdd(&acc->branched[lane], s_br[lane]);
                atomicAdd(&acc->evals[lane], s_ev[lane]);
            }
            if (lane == 0) {
                if (s_children) atomicAdd(&acc->children, s_children);
                if (s_overflow) atomicOr(&acc->overflow, 1u);
                mine->children = s_children;
                __threadfence();
                *(volatile unsigned*)&mine->tag = tag;
            }
            unsigned base = 0;
            for (int c = (int)lane; c < b; c += 32) {
                const PersistTotals* t = &scratch->totals[par][c];
                while (*(volatile const unsigned*)&t->tag != tag) { }
                __threadfence();
                base += __ldcg(&t->children);
            }
            base = __reduce_add_sync(kFullMask, base);
            if (lane == 0) s_base = base;
        }
        __syncthreads();
        if (clocked) { const unsigned long long t = global_ns(); tp1 += t - tmark; tmark = t; }
        // ---- emit: next node list, tree entries; K = 1: the -q test at every node
        {
            const unsigned base = s_base;
            if (K == 1) {
                bool stop = false;
                for (int i = (int)tid; i < 2 * nb; i += (int)blockDim.x) {
                    const int k = i >> 1, br = i & 1, j = j0 + k;
                    const unsigned open = s_cnt[k];
                    unsigned pos = base + s_off[k];
                    if (br == 0 && Q > 0 && (unsigned)(m - j) + pos >= (unsigned)Q) stop = true;   // NDP:894 before popping node j
                    if (!(open & (1u << br))) continue;
                    if (br == 1) pos += open & 1u;
                    const BfsChild c = children[2 * j + br];
                    const int4 nraw = __ldcg(reinterpret_cast<const int4*>(&nodes[j]));
                    const int var = nraw.z == K_UNIT ? (nraw.w < 0 ? -nraw.w : nraw.w) : nraw.w;
                    BfsNode nn;
                    nn.slot = (uint32_t)(2 * j + br); nn.n = c.n; nn.kind = c.kind; nn.lit = c.lit;
                    nnodes[pos] = nn;
                    nntree[pos] = (int32_t)(tree_count + pos);
                    tree_parent[tree_count + pos] = __ldcg(&ntree[j]);
                    tree_lit[tree_count + pos] = br ? -var : var;                               // LA, RA (NDP:904-925)
                }
                if (stop) *(volatile unsigned*)&scratch->stop_tag[par] = tag;
            } else {
                // one thread per (node, output, relative level): K chained tree entries per output
                const int per_node = fan * K;
                for (int i = (int)tid; i < nb * per_node; i += (int)blockDim.x) {
                    const int k = i / per_node, rem = i - k * per_node, e = rem / K, r = rem - e * K, j = j0 + k;
                    if (e >= (int)s_cnt[k]) continue;
                    const unsigned pos = base + s_off[k] + (unsigned)e;
                    const size_t o = (size_t)fan * j + e;
                    const unsigned long long t0 = tree_count + (unsigned long long)pos * K;
                    tree_parent[t0 + r] = r == 0 ? __ldcg(&ntree[j]) : (int32_t)(t0 + r - 1);
                    tree_lit[t0 + r] = (int32_t)paths[o * kWindowMax + r];
                    if (r == K - 1) {
                        const BfsChild c = children[o];
                        BfsNode nn;
                        nn.slot = (uint32_t)o; nn.n = c.n; nn.kind = c.kind; nn.lit = c.lit;
                        nnodes[pos] = nn;
                        nntree[pos] = (int32_t)(t0 + K - 1);
                    }
                }
            }
        }
        __threadfence();
        if (clocked) { const unsigned long long t = global_ns(); tp2 += t - tmark; tmark = t; }
        grid.sync();
        if (clocked) { const unsigned long long t = global_ns(); tp3 += t - tmark; tmark = t; }
        // ---- commit: every CTA reads the same grid-wide sums
        if (tid < (unsigned)K) {
            const PersistAcc* acc = &scratch->acc[step % 3];
            t_exp[tid] = __ldcg(&acc->expanded[tid]);
            t_push[tid] = __ldcg(&acc->pushed[tid]);
            t_br[tid] = __ldcg(&acc->branched[tid]);
            t_ev[tid] = __ldcg(&acc->evals[tid]);
            if (tid == 0) {
                t_children = __ldcg(&acc->children);
                t_overflow = __ldcg(&acc->overflow);
                s_stop = __ldcg(&scratch->stop_tag[par]) == tag ? 1u : 0u;
            }
        }
        if (b == 0 && tid >= 32) {
            // clear the accumulator of step + 2: last read before this step's barrier, next written after the next one
            for (unsigned i = tid - 32; i < sizeof(PersistAcc) / 4; i += blockDim.x - 32)
                reinterpret_cast<unsigned*>(&scratch->acc[(step + 2) % 3])[i] = 0u;
        }
        __syncthreads();
        if (K == 1) {
            if (s_stop) { reason = PEXIT_HOST_LEVEL; break; }   // a -q stop fires inside this level: leave it to the host
            it += (long long)t_exp[0];                          // NDP:926
            task_count += (long long)t_children;                // NDP:912,923
            evals_total += t_ev[0];
            tree_count += t_children;
            m = (int)t_children;
            cur ^= 1;
            cur_buf ^= 1;
            levels_done += 1;
            if (!near_stop && Kcur < Kmax) Kcur = min(Kmax, Kcur * 2);   // regrow after a fan-out overflow
        } else {
            // How many of the window's levels are certain to have run without any stop test firing.
            // -q (NDP:894): the queue holds m_r entries when level r starts and moves by (children - 1)
            // per pop, so it never exceeds m_r + (nodes of the level with two open children).
            // -d (NDP:935): the level is consumed whole iff iterations + m_r < D.
            int valid = 0;
            long long it_r = it;
            long long m_r = m;
            for (int r = 0; r < K; ++r) {
                const bool ok = m_r == 0 || (Q > 0 ? m_r + (long long)t_br[r] < (long long)Q : it_r + m_r < (long long)D);
                if (!ok) break;
                it_r += (long long)t_exp[r];
                m_r = (long long)t_push[r];
                ++valid;
            }
            if (valid == K && !t_overflow) {
                long long pushes = 0;
                for (int r = 0; r < K; ++r) { pushes += (long long)t_push[r]; evals_total += t_ev[r]; }
                it = it_r;
                task_count += pushes;
                tree_count += (unsigned long long)t_children * K;
                m = (int)t_children;
                cur ^= 1;
                cur_buf ^= 1;
                levels_done += K;
                ++windows_ok;
                if (!near_stop) Kcur = min(Kmax, Kcur * 2);
            } else {
                // redo from the untouched source level with a shorter window
                ++windows_redone;
                if (valid < K) { near_stop = true; Kcur = max(1, valid); }
                if (t_overflow) Kcur = max(1, min(Kcur, K / 2));
            }
        }
        __syncthreads();   // the shared totals are rewritten by the next step
        if (clocked) { const unsigned long long t = global_ns(); tp4 += t - tmark; tmark = t; }
    }
    if (b == 0 && tid == 0) {
        st->m = m; st->cur = cur; st->cur_buf = cur_buf; st->levels_done = levels_done; st->exit_reason = reason;
        st->it = it; st->task_count = task_count; st->evals = evals_total; st->tree_count = tree_count;
        st->windows_ok = windows_ok; st->windows_redone = windows_redone; st->window = Kcur;
        st->t_phase[0] = tp0; st->t_phase[1] = tp1; st->t_phase[2] = tp2; st->t_phase[3] = tp3; st->t_phase[4] = tp4;
    }
}

// State slots from assignment tables: one warp per slot.  tables[q] (table_words each) is copied
// into the slot, the bitmaps are derived from it, and n/kind/lit are reported.  Used for the root
// (empty table) and for frontiers that were not produced by the index BFS.
__global__ void __launch_bounds__(128) index_states_kernel(const IndexRoot ix, const unsigned* __restrict__ tables,
                                                           unsigned* __restrict__ slots, int count,
                                                           BfsNode* __restrict__ meta) {
    extern __shared__ __align__(16) unsigned smem_words[];
    const unsigned lane = threadIdx.x & 31u;
    const int warps_per_block = blockDim.x >> 5;
    unsigned* S = smem_words + (size_t)(threadIdx.x >> 5) * ix.slot_words;
    unsigned* table = S + ix.table_off;
    for (int q = blockIdx.x * warps_per_block + (threadIdx.x >> 5); q < count; q += gridDim.x * warps_per_block) {
        __syncwarp();
        for (int w = lane; w < ix.slot_words - ix.table_off; w += 32)
            table[w] = w < ix.table_words ? __ldg(&tables[(size_t)q * ix.table_words + w]) : 0u;
        __syncwarp();
        const int n_live = index_rebuild(ix, S, table, lane);
        int kind, lit;
        index_choice(ix, S, table, n_live, lane, kind, lit);
        slot_copy(slots + (size_t)q * ix.slot_words, S, ix.slot_words, lane);
        if (lane == 0) { BfsNode nn; nn.slot = (uint32_t)q; nn.n = n_live; nn.kind = kind; nn.lit = lit; meta[q] = nn; }
    }
}

// Materialise clause lists: out[q] (stride records, tombstone-padded tail) = root filtered by the
// table of state slot_ptr[q], in root order.  One warp per element.
__global__ void __launch_bounds__(128) index_materialise_kernel(const IndexRoot ix,
                                                                const unsigned* const* __restrict__ slot_ptr,
                                                                int count, pclause* __restrict__ out, size_t stride,
                                                                int32_t* __restrict__ out_n) {
    const unsigned lane = threadIdx.x & 31u;
    const int warps_per_block = blockDim.x >> 5;
    for (int q = blockIdx.x * warps_per_block + (threadIdx.x >> 5); q < count; q += gridDim.x * warps_per_block) {
        const unsigned* table = slot_ptr[q] + ix.table_off;
        SetInfo s = warp_rebuild<true>(ix.root, ix.n_root, out + (size_t)q * stride, table, lane);
        if (lane == 0 && out_n) out_n[q] = s.n;
    }
}

// Does list a[q] (n[q] records at a + q * stride_a) equal list b[q]?  One warp per element; any difference
// (a length that is not expect_n[q] included) raises *mismatch.  Used to check that a resumed frontier
// really is "the root filtered by its choices" before it is handed to the index engine.
__global__ void __launch_bounds__(128) clause_lists_equal_kernel(const pclause* __restrict__ a, size_t stride_a,
                                                                 const int32_t* __restrict__ n_a,
                                                                 const pclause* __restrict__ b, size_t stride_b,
                                                                 const int32_t* __restrict__ expect_n, int count,
                                                                 unsigned* __restrict__ mismatch) {
    const unsigned lane = threadIdx.x & 31u;
    const int warps_per_block = blockDim.x >> 5;
    for (int q = blockIdx.x * warps_per_block + (threadIdx.x >> 5); q < count; q += gridDim.x * warps_per_block) {
        const int n = n_a[q];
        bool bad = n != expect_n[q];
        if (!bad) {
            const pclause* pa = a + (size_t)q * stride_a;
            const pclause* pb = b + (size_t)q * stride_b;
            for (int i = lane; i < n; i += 32) {
                const pclause x = pa[i], y = pb[i];
                // literal codes only: the cached zero count is derived data
                bad |= x.x != y.x || (x.y & 0xffffu) != (y.y & 0xffffu);
            }
        }
        if (__any_sync(kFullMask, bad) && lane == 0) atomicOr(mismatch, 1u);
    }
}

// ------------------------------------------------------------------------------------------
// DFS over state slots.  Persistent grid; a warp pulls the next subproblem of its shard, copies its
// state slot into shared memory and runs the whole search there.  The explicit stack is the trail
// (global memory) plus one pending bit per trail position; a popped LA sibling is re-derived from
// (slot table + trail) with one pass over the root.
// The unit of work is a PIECE: a whole subproblem of the shard (orig == nullptr: piece q is
// subproblem q), or -- after the DFS-order split of ndp_split.cuh -- a subtree of one, listed in
// the order Satisfy_iterative would reach it.  best_piece[o] is the lowest piece of subproblem o
// that holds a model: later pieces of o are abandoned, because the reference stops a subproblem
// at its first model (NDP:823,849,869).
constexpr int K_LEAF = 3;   // piece kind: a branch the split found empty -- the model itself, no node left to visit

struct IndexParams {
    IndexRoot ix;
    const unsigned* const* slot_ptr;  // [n_local] state slot of each piece
    const BfsNode* meta;              // [n_local] n / kind / lit of each piece
    const int32_t* orig;              // [n_local] subproblem (local index) a piece belongs to, or nullptr = identity
    unsigned int* best_piece;         // [subproblems of the shard] lowest piece with a model (0xffffffff: none)
    int pend_words, trail_cap;
    unsigned long long first, stride, n_local;   // n_local = pieces
    unsigned int* next;
    unsigned int* progress;           // host-mapped word: last unit pulled (ndp_dfs_progress), or nullptr
    unsigned long long* best;
    unsigned long long* peer_best[8];
    int n_peers;
    int early_exit;
    uint8_t* verdict;
    unsigned long long* nodes;
    unsigned long long* evals;
    unsigned int* rebuilds;
    unsigned long long* model_sub;    // [n_warps] PIECE whose model the warp parked
    int32_t* model_len;               // [n_warps]
    unsigned* snaps;                  // [n_warps][snap_depth] state slots saved at decisions with a pending LA sibling
    int snap_depth;                   //   (0: always re-derive the sibling from the trail)
    short* trails;                    // [n_warps][trail_cap] working trail of each warp
    int32_t* models;                  // [n_warps][trail_cap] trail parked by the warp holding the lowest model
    int warp_bytes, off_pend;         // shared memory of one searcher: state slot, then pend
    unsigned char* gstate;            // [searchers][warp_bytes] the same in global memory (lane-group kernel, kGlobal)
    unsigned max_steps;               // profiling knob (0 = off): a warp gives up after this many unit steps (pieces left NOT_RUN)
};

// kMinBlocks: resident CTAs per SM the register allocation must allow (8: 64 registers, 32 warps per SM;
// 12: 40 registers, 48 warps per SM -- for state slots small enough that shared memory holds 48 of them).
template <bool kPlain, int kMinBlocks = 8>
__global__ void __launch_bounds__(128, kMinBlocks) dfs_index_kernel(const __grid_constant__ IndexParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const IndexRoot& ix = p.ix;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned char* mine = smem_raw + (size_t)warp * p.warp_bytes;
    unsigned* B = reinterpret_cast<unsigned*>(mine);
    unsigned* table = B + ix.table_off;
    unsigned* pend = reinterpret_cast<unsigned*>(mine + p.off_pend);
    const unsigned tbase = gwarp * (unsigned)p.trail_cap;   // this warp's trail: p.trails[tbase + i] (32-bit index arithmetic)
    unsigned* snap = p.snaps + (size_t)gwarp * p.snap_depth * ix.slot_words;

    for (;;) {
        unsigned q = 0;
        if (lane == 0) {
            q = atomicAdd(p.next, 1u);
            if (p.progress) *(volatile unsigned*)p.progress = q;   // posted write to pinned host memory, once per unit
        }
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= p.n_local) break;
        const unsigned o = p.orig ? (unsigned)p.orig[q] : q;
        const unsigned long long g = p.first + (unsigned long long)o * p.stride;
        volatile unsigned int* my_best = p.best_piece + o;
        if ((p.early_exit && *(volatile unsigned long long*)p.best < g) || *my_best < q) {
            if (lane == 0) { p.verdict[q] = 2; p.nodes[q] = 0; p.evals[q] = 0; p.rebuilds[q] = 0; }
            continue;
        }
        const unsigned* slot = p.slot_ptr[q];
        const BfsNode nd = p.meta[q];
        int n_live = nd.n, kind = nd.kind, lit = nd.lit;
        int ntrail = 0;
        unsigned long long nodes = 0, evals = 0;
        unsigned rebuilds = 0;
        int verdict = 0;
        int npend = 0, snap_low = 0, snap_n = 0;   // pending siblings; lowest one whose snapshot is intact; lane k: |A| of snapshot k
        __syncwarp();
        if (kind != K_LEAF) {
            slot_copy(B, slot, ix.slot_words, lane);
            for (int w = lane; w < p.pend_words; w += 32) pend[w] = 0u;
        }
        __syncwarp();
        if (kind == K_LEAF) verdict = 1;   // the split already reached the model: no node left to visit
        else if (n_live == 0 || lit == 0) { nodes = 1; verdict = 1; }   // choice()==0 at the piece's own root (NDP:817-823);
        else for (;;) {                                                 // below it every node has a literal to branch on
            ++nodes;  // one Satisfy_iterative loop iteration (NDP:801-806)
            if ((nodes & 63u) == 0 &&
                ((p.early_exit && *(volatile unsigned long long*)p.best < g) || *my_best < q)) { verdict = 2; break; }
            const int var = lit < 0 ? -lit : lit;                  // K_DECIDE carries the variable itself
            evals += (unsigned long long)n_live;                   // one ResolutionStepWithConflict (NDP:831)
            int t = lit;   // unit clause: the other branch falsifies it ({0,0,0}, never pushed)
            if (kind != K_UNIT) {
                bool la_conf, ra_conf;
                int la_n, ra_n;
                index_classify(ix, B, n_live, var, la_n, la_conf, ra_n, ra_conf, lane);
                if (la_n == 0) {                                    // LA empty wins first (NDP:843-850)
                    if (lane == 0) p.trails[tbase + (unsigned)(ntrail)] = (short)var;
                    ++ntrail; verdict = 1; break;
                }
                if (ra_n == 0) {                                    // NDP:863-870
                    if (lane == 0) p.trails[tbase + (unsigned)(ntrail)] = (short)(-var);
                    ++ntrail; verdict = 1; break;
                }
                t = !ra_conf ? -var : (!la_conf ? var : 0);         // RA is explored first
                if (!ra_conf && !la_conf) {
                    // both open: LA stays pending.  Keep the state before the branch: the LA sibling is one
                    // resolution step away from it
                    if (p.snap_depth > 0) {
                        const int k = npend % p.snap_depth;
                        __syncwarp();
                        slot_copy(snap + (size_t)k * ix.slot_words, B, ix.slot_words, lane);
                        if ((int)lane == k) snap_n = n_live;
                        snap_low = max(min(snap_low, npend), npend - p.snap_depth + 1);
                    }
                    ++npend;
                    if (lane == 0) pend[ntrail >> 5] |= 1u << (ntrail & 31);
                }
            }
            if (t != 0) {
                // every lane stores the same value to the same place: no divergent lane-0 block on the hot path
                p.trails[tbase + (unsigned)(ntrail)] = (short)t;
                table_set(table, t);
                ++ntrail;
                bool conflict;
                n_live = index_assign<true, kPlain>(ix, B, n_live, t, conflict, lane);
                if (n_live == 0) { verdict = 1; break; }            // empty branch => model (NDP:843-852 / 863-872)
                if (!conflict) {
                    index_choice<kPlain>(ix, B, table, n_live, lane, kind, lit);
                    if (!kPlain && lit == 0) { ++nodes; verdict = 1; break; }   // choice()==0 (only with a {0,0,0} root clause)
                    continue;
                }
            }
            // Backtrack: pop to the deepest decision whose LA sibling is still pending.
            __syncwarp();
            int pos = -1;
            for (int w = (ntrail - 1) >> 5; w >= 0 && ntrail > 0; --w) {
                unsigned bits = pend[w];
                if (w == ((ntrail - 1) >> 5)) {
                    const int top = (ntrail - 1) & 31;
                    if (top != 31) bits &= (2u << top) - 1u;
                }
                if (bits) { pos = w * 32 + 31 - __clz(bits); break; }
            }
            if (pos < 0) break;   // stack empty: UNSAT
            __syncwarp();
            if (lane == 0) {
                pend[pos >> 5] &= ~(1u << (pos & 31));
                p.trails[tbase + (unsigned)(pos)] = (short)(-p.trails[tbase + (unsigned)(pos)]);   // -i -> +i : now the LA sibling
            }
            ntrail = pos + 1;
            --npend;
            ++rebuilds;
            if (p.snap_depth > 0 && npend >= snap_low) {
                const int k = npend % p.snap_depth;
                __syncwarp();
                __threadfence_block();
                slot_copy(B, snap + (size_t)k * ix.slot_words, ix.slot_words, lane);
                const int la = (int)p.trails[tbase + (unsigned)(pos)];
                const int n_before = __shfl_sync(kFullMask, snap_n, k);
                __syncwarp();
                if (lane == 0) table_set(table, la);
                bool conflict;
                n_live = index_assign<true, kPlain>(ix, B, n_before, la, conflict, lane);   // open and non-empty: it was left pending
                index_choice<kPlain>(ix, B, table, n_live, lane, kind, lit);
                continue;
            }
            snap_low = min(snap_low, npend);   // siblings pushed from here on have intact snapshots again
            for (int w = lane; w < ix.table_words; w += 32) table[w] = __ldg(&slot[ix.table_off + w]);
            __syncwarp();
            __threadfence_block();
            for (int j = lane; j < ntrail; j += 32) {
                const int l = p.trails[tbase + (unsigned)(j)];
                const unsigned v = (unsigned)(l < 0 ? -l : l);
                atomicOr(&table[v >> 4], (l > 0 ? 1u : 2u) << ((v & 15u) * 2u));
            }
            __syncwarp();
            n_live = index_rebuild(ix, B, table, lane);
            index_choice<kPlain>(ix, B, table, n_live, lane, kind, lit);
        }
        if (verdict == 1) {
            int park = 0;
            if (lane == 0) {
                const unsigned old_piece = atomicMin(p.best_piece + o, q);
                const unsigned long long old = atomicMin(p.best, g);
                if (g < old)
                    {
#pragma unroll
                        for (int d = 0; d < 8; ++d)   // constant indices: the parameter block stays in constant memory
                            if (d < p.n_peers) atomicMin_system(p.peer_best[d], g);
                    }
                park = q < old_piece && g <= old;
            }
            park = __shfl_sync(kFullMask, park, 0);
            if (park) {   // so far the first model of the lowest SAT subproblem: park it in this warp's slot
                __syncwarp();
                __threadfence_block();
                int32_t* out = p.models + (size_t)gwarp * p.trail_cap;
                for (int j = lane; j < ntrail; j += 32) out[j] = (int32_t)p.trails[tbase + (unsigned)(j)];
                if (lane == 0) { p.model_sub[gwarp] = q; p.model_len[gwarp] = ntrail; }
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

// One thread per subproblem: walk its choices up the BFS tree and set them in its table row.
__global__ void prefix_table_kernel(const int32_t* __restrict__ entry_tree, const int32_t* __restrict__ tree_parent,
                                    const int32_t* __restrict__ tree_lit, unsigned long long first,
                                    unsigned long long stride, unsigned long long n_local, int table_words,
                                    unsigned* __restrict__ tables) {
    for (unsigned long long q = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; q < n_local;
         q += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned* row = tables + q * table_words;
        for (int32_t t = entry_tree[first + q * stride]; t > 0; t = tree_parent[t]) {
            const int l = tree_lit[t];
            const unsigned v = (unsigned)(l < 0 ? -l : l);
            row[v >> 4] |= (l > 0 ? 1u : 2u) << ((v & 15u) * 2u);
        }
    }
}

// Gather scattered state slots into a dense buffer (slot q <- ptr[q]); one warp per slot.
__global__ void __launch_bounds__(128) slot_gather_kernel(const unsigned* const* __restrict__ ptr, int count,
                                                          int slot_words, unsigned* __restrict__ dst) {
    const unsigned lane = threadIdx.x & 31u;
    const int warps_per_block = blockDim.x >> 5;
    for (int q = blockIdx.x * warps_per_block + (threadIdx.x >> 5); q < count; q += gridDim.x * warps_per_block)
        slot_copy(dst + (size_t)q * slot_words, ptr[q], slot_words, lane);
}

// One thread: the choices of one subproblem, leaf to root (the host reverses them).
__global__ void tree_path_kernel(const int32_t* __restrict__ tree_parent, const int32_t* __restrict__ tree_lit,
                                 int32_t leaf, int32_t* __restrict__ out, int cap, int32_t* __restrict__ out_len) {
    int n = 0;
    for (int32_t t = leaf; t > 0; t = tree_parent[t]) {
        if (n < cap) out[n] = tree_lit[t];
        ++n;
    }
    *out_len = n;
}

}  // namespace ndp
