// ndp_clause.cuh -- packed clause representation and the warp-wide resolution passes.
//
// A clause is the reference's Clause3 (ClauseSetPool.hpp:37-39: three int literals, 0 = "already
// false / padding") packed into 8 bytes so that one lane moves one clause with a single 64-bit
// access and a warp moves 32 clauses (256 B) per instruction:
//
//     x = c0 | c1 << 16          y = c2 | zc << 16
//
// c_j is the 16-bit sign-magnitude CODE of literal l_j:  0 for literal 0, else (|l| << 1) | (l < 0).
// With this code one XOR against code(t) answers both questions the resolution step asks of a
// literal (NDP-5_6_7.cpp:671-680): result 0 => the literal IS t (clause satisfied), result 1 =>
// the literal is -t (it becomes 0).  zc caches the clause's zero count, which is what choice()
// (NDP-5_6_7.cpp:753-781) keys on.  A record with x == y == 0xffffffff is a TOMBSTONE: a clause
// that was satisfied and dropped but not yet squeezed out (lazy deletion, DFS only).
//
// Clause sets are dense arrays of these records ("slots" of a device arena).  Stable order is
// preserved everywhere, because choice() is first-in-order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ndp {

typedef uint2 pclause;

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kMaxVar = 32766;   // code 0xfffe/0xffff (variable 32767) is reserved for tombstones
constexpr unsigned kTomb = 0xffffffffu;

enum : int { K_EMPTY = 0, K_UNIT = 1, K_DECIDE = 2 };

// What choice() yields for a clause set, plus how it got there:
//  K_EMPTY : the set is empty (choice() == 0, NDP:780)
//  K_UNIT  : a clause with two zeros exists; lit = its (signed) remaining literal (NDP:756-766)
//  K_DECIDE: lit = the variable (>= 0) from the first one-zero clause's last literal
//            (NDP:767-777) or |A[0].l[0]| (NDP:778-779); lit == 0 only when A[0] is {0,0,0}.
struct SetInfo {
    int n;         // live clauses in the set
    int conflict;  // some clause is {0,0,0}
    int kind;
    int lit;
};

__host__ __device__ __forceinline__ unsigned lit_code(int l) {
    return l == 0 ? 0u : (((unsigned)(l < 0 ? -l : l)) << 1) | (l < 0 ? 1u : 0u);
}
__host__ __device__ __forceinline__ int code_lit(unsigned c) {
    int v = (int)(c >> 1);
    return (c & 1u) ? -v : v;
}
__host__ __device__ __forceinline__ pclause pack_clause(int l0, int l1, int l2) {
    pclause c;
    unsigned zc = (l0 == 0) + (l1 == 0) + (l2 == 0);
    c.x = lit_code(l0) | (lit_code(l1) << 16);
    c.y = lit_code(l2) | (zc << 16);
    return c;
}
__host__ __device__ __forceinline__ void unpack_clause(pclause c, int& l0, int& l1, int& l2) {
    l0 = code_lit(c.x & 0xffffu);
    l1 = code_lit(c.x >> 16);
    l2 = code_lit(c.y & 0xffffu);
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Per-lane resolution of one clause against "literal with code ct becomes true".
// Returns true when the clause mentions the variable at all ("touched").
//   sat  : the clause holds the true literal (dropped in this branch, NDP:671-673 / 674-676)
//   else : occurrences of the false literal are zeroed and zc is bumped; x,y updated in place.
__device__ __forceinline__ bool resolve_clause(unsigned& x, unsigned& y, unsigned t2, unsigned ct, bool& sat) {
    const unsigned zx = x ^ t2, zy = y ^ ct;
    const bool m0 = (zx & 0x0000fffeu) == 0u, m1 = (zx & 0xfffe0000u) == 0u, m2 = (zy & 0x0000fffeu) == 0u;
    sat = false;
    if (!(m0 | m1 | m2)) return false;
    sat = (m0 && !(zx & 1u)) || (m1 && !(zx & 0x10000u)) || (m2 && !(zy & 1u));
    if (!sat) {
        if (m0) x &= 0xffff0000u;
        if (m1) x &= 0x0000ffffu;
        if (m2) y &= 0xffff0000u;
        y += ((unsigned)m0 + (unsigned)m1 + (unsigned)m2) << 16;
    }
    return true;
}
// cheap "does this clause mention the variable" test for the streaming phase
__device__ __forceinline__ bool touches(unsigned x, unsigned y, unsigned t2, unsigned ct) {
    const unsigned vx = (x ^ t2) & 0xfffefffeu;
    const unsigned hz = (vx - 0x00010001u) & ~vx & 0x80008000u;   // a zero halfword in vx
    return (hz != 0u) | (((y ^ ct) & 0x0000fffeu) == 0u);
}

// Running answer to "what would choice() return for the live clauses seen so far".
struct ChoiceTracker {
    int lit2 = 0;        // literal of the first live two-zero clause (0 = none yet)
    int lit1 = 0;        // last literal of the first live one-zero clause (0 = none yet)
    int first = 0;       // l[0] of the first live clause
    bool have_first = false;

    // x,y: this lane's post-resolution record; live: the lane holds a live clause.
    __device__ __forceinline__ void update(bool live, unsigned x, unsigned y) {
        const unsigned zc = (y >> 16) & 3u;
        if (!have_first) {
            unsigned ma = __ballot_sync(kFullMask, live);
            if (ma) { first = code_lit(__shfl_sync(kFullMask, x & 0xffffu, __ffs(ma) - 1)); have_first = true; }
        }
        unsigned m2 = __ballot_sync(kFullMask, live && zc == 2u);
        if (m2) {
            unsigned only = (x & 0xffffu) | (x >> 16) | (y & 0xffffu);
            lit2 = code_lit(__shfl_sync(kFullMask, only, __ffs(m2) - 1));
        } else if (lit1 == 0) {
            unsigned m1 = __ballot_sync(kFullMask, live && zc == 1u);
            if (m1) {
                unsigned last = (y & 0xffffu) ? (y & 0xffffu) : (x >> 16);
                lit1 = code_lit(__shfl_sync(kFullMask, last, __ffs(m1) - 1));
            }
        }
    }
    __device__ __forceinline__ void finish(SetInfo& s) const {
        if (s.n == 0) { s.kind = K_EMPTY; s.lit = 0; }
        else if (lit2) { s.kind = K_UNIT; s.lit = lit2; }
        else { s.kind = K_DECIDE; int v = lit1 ? lit1 : first; s.lit = v < 0 ? -v : v; }
    }
};

// ------------------------------------------------------------------------------------------
// EAGER pass (BFS): one branch of the resolution step over src[0..n) written, stably compacted,
// to dst (a different slot).  ResolutionStep + the conflict scan + the child's choice() in one
// read of the parent (NDP-5_6_7.cpp:700-750, 905-907, 753-781).  src holds no tombstones.
__device__ __forceinline__ SetInfo warp_apply_copy(const pclause* __restrict__ src, int n, pclause* __restrict__ dst,
                                                   int t, unsigned lane) {
    const unsigned ct = lit_code(t), t2 = ct | (ct << 16);
    const unsigned lt = lanemask_lt();
    ChoiceTracker trk;
    int out = 0;
    bool confl = false;
    constexpr int kU = 8;   // global loads in flight per lane: the pass is latency-bound on narrow levels
    for (int base = 0; base < n; base += 32 * kU) {
        pclause cc[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int idx = base + u * 32 + (int)lane;
            cc[u] = idx < n ? __ldg(&src[idx]) : make_uint2(kTomb, kTomb);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (base + u * 32 < n) {   // warp-uniform
                pclause c = cc[u];
                const bool valid = c.x != kTomb;   // src holds no tombstones; padding lanes do
                bool sat;
                resolve_clause(c.x, c.y, t2, ct, sat);
                const bool emit = valid && !sat;
                confl |= emit && ((c.y >> 16) & 3u) == 3u;
                const unsigned emask = __ballot_sync(kFullMask, emit);
                if (trk.lit2 == 0) trk.update(emit, c.x, c.y);
                if (emit) dst[out + __popc(emask & lt)] = c;
                out += __popc(emask);
            }
        }
    }
    SetInfo s;
    s.n = out;
    s.conflict = __any_sync(kFullMask, confl) ? 1 : 0;
    trk.finish(s);
    return s;
}

// ------------------------------------------------------------------------------------------
// LAZY passes (DFS).  The working set W[0..n_buf) is updated IN PLACE: satisfied clauses become
// tombstones instead of being squeezed out, and only clauses that mention the variable are
// written (a handful per pass), so a pass is a read-only stream over shared memory.
// Invariant kept by warp_rebuild / warp_compact: W is padded with tombstones up to the next
// multiple of kGroup records, so the streaming loops need no bounds checks.
constexpr int kStreamUnroll = 4;
constexpr int kGroup = 32 * kStreamUnroll;
__host__ __device__ __forceinline__ int pad_up(int n) { return (n + kGroup - 1) / kGroup * kGroup; }

__device__ __forceinline__ void pad_tail(pclause* W, int n, unsigned lane) {
    for (int idx = n + (int)lane; idx < pad_up(n); idx += 32) W[idx] = make_uint2(kTomb, kTomb);
}

// A touched clause, parked in registers until the lane meets its next one (or the pass ends):
// the stream itself stays free of the (divergent) resolution code.
struct Parked {
    int idx = -1;
    unsigned x = 0, y = 0;
};
template <bool kWrite>
__device__ __forceinline__ void settle(pclause* W, Parked& p, unsigned t2, unsigned ct, int& dropped, bool& confl) {
    bool sat;
    resolve_clause(p.x, p.y, t2, ct, sat);
    if (sat) { p.x = kTomb; p.y = kTomb; ++dropped; }
    else confl |= ((p.y >> 16) & 3u) == 3u;
    if (kWrite) W[p.idx] = make_uint2(p.x, p.y);
    p.idx = -1;
}

// One branch of the resolution step ("literal t becomes true") over the whole set.
//   kWrite = true : apply it (ResolutionStepWithConflict for the branch that is explored,
//                   NDP-5_6_7.cpp:654-697) and report what choice() will return next:
//       phase A  chunk by chunk with the ChoiceTracker, until the next unit clause is known;
//       phase B  the rest of the buffer, kStreamUnroll chunks per step, no warp collectives.
//   kWrite = false: classify only (decision nodes look at both children before committing).
// n_live = live clauses on entry; sticky_conflict = the set already holds a {0,0,0} clause
// (it can never leave).  Result: s.n = live clauses afterwards, s.conflict, s.kind/lit (kWrite).
template <bool kWrite>
__device__ __forceinline__ SetInfo warp_pass_lazy(pclause* W, int n_buf, int n_live, int t, bool sticky_conflict,
                                                  unsigned lane) {
    const unsigned ct = lit_code(t), t2 = ct | (ct << 16);
    const int n_pad = pad_up(n_buf);
    ChoiceTracker trk;
    int dropped = 0;
    bool confl = false;
    int base = 0;
    if (kWrite) {
        // ---- phase A: in order, with the tracker, until the first live two-zero clause shows up
        for (; base < n_pad && trk.lit2 == 0; base += 32) {
            const int idx = base + (int)lane;
            pclause c = W[idx];
            bool live = c.x != kTomb;
            const unsigned zc = (c.y >> 16) & 3u;
            const bool touched = touches(c.x, c.y, t2, ct);
            // can this chunk change the tracker or the set at all?
            const bool interesting = live && (touched || zc == 2u || (trk.lit1 == 0 && zc == 1u) || !trk.have_first);
            if (__ballot_sync(kFullMask, interesting) == 0u) continue;
            if (touched) {
                bool sat;
                resolve_clause(c.x, c.y, t2, ct, sat);
                if (sat) { c.x = kTomb; c.y = kTomb; ++dropped; live = false; }
                else confl |= ((c.y >> 16) & 3u) == 3u;
                W[idx] = c;
            }
            trk.update(live, c.x, c.y);
        }
        // finish the current group chunk by chunk so that phase B starts group-aligned
        for (; base < n_pad && (base % kGroup) != 0; base += 32) {
            const int idx = base + (int)lane;
            pclause c = W[idx];
            if (touches(c.x, c.y, t2, ct)) {
                Parked p;
                p.idx = idx; p.x = c.x; p.y = c.y;
                settle<true>(W, p, t2, ct, dropped, confl);
            }
        }
    }
    // ---- phase B: pure stream
    Parked park;
    for (; base < n_pad; base += kGroup) {
        pclause c[kStreamUnroll];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) c[u] = W[base + u * 32 + (int)lane];
        static_assert(kStreamUnroll == 4, "the select chain below is written for 4 chunks per group");
        const bool h0 = touches(c[0].x, c[0].y, t2, ct), h1 = touches(c[1].x, c[1].y, t2, ct);
        const bool h2 = touches(c[2].x, c[2].y, t2, ct), h3 = touches(c[3].x, c[3].y, t2, ct);
        if (!(h0 | h1 | h2 | h3)) continue;
        unsigned hit = (h0 ? 1u : 0u) | (h1 ? 2u : 0u) | (h2 ? 4u : 0u) | (h3 ? 8u : 0u);
        while (hit) {   // rare and divergent: a lane that meets a clause of the variable parks it
            const int u = __ffs(hit) - 1;
            hit &= hit - 1u;
            if (park.idx >= 0) settle<kWrite>(W, park, t2, ct, dropped, confl);
            park.idx = base + u * 32 + (int)lane;
            park.x = u == 0 ? c[0].x : u == 1 ? c[1].x : u == 2 ? c[2].x : c[3].x;
            park.y = u == 0 ? c[0].y : u == 1 ? c[1].y : u == 2 ? c[2].y : c[3].y;
        }
    }
    if (__any_sync(kFullMask, park.idx >= 0)) {       // one convergent round for the parked clauses
        if (park.idx >= 0) settle<kWrite>(W, park, t2, ct, dropped, confl);
    }
    SetInfo s;
    s.n = n_live - (int)__reduce_add_sync(kFullMask, (unsigned)dropped);
    s.conflict = (sticky_conflict || __any_sync(kFullMask, confl)) ? 1 : 0;
    s.kind = K_EMPTY;
    s.lit = 0;
    if (kWrite) trk.finish(s);
    return s;
}

// Squeeze the tombstones out of W[0..n_buf) in place (stable) and re-pad.  Returns the live count.
__device__ __noinline__ int warp_compact(pclause* W, int n_buf, unsigned lane) {
    const unsigned lt = lanemask_lt();
    int out = 0;
    for (int base = 0; base < n_buf; base += 32) {
        const int idx = base + (int)lane;
        pclause c = idx < n_buf ? W[idx] : make_uint2(kTomb, kTomb);
        const bool live = c.x != kTomb;
        const unsigned m = __ballot_sync(kFullMask, live);
        if (live && out + __popc(m & lt) != idx) W[out + __popc(m & lt)] = c;
        out += __popc(m);
    }
    __syncwarp();
    pad_tail(W, out, lane);
    return out;
}

// 2-bit-per-variable assignment table: 0 unassigned, 1 true, 2 false.
__device__ __forceinline__ unsigned table_get(const unsigned* table, unsigned var) {
    return (table[var >> 4] >> ((var & 15u) * 2u)) & 3u;
}

// Re-materialise a clause set from its base slot under the assignment in `table`
// (the stable-filter equivalence of SURVEY.md §8a: a node's set == its base filtered by the
// assignment).  With kTable == false it is a plain copy that also computes the SetInfo.
template <bool kTable>
__device__ __noinline__ SetInfo warp_rebuild(const pclause* __restrict__ base, int nb, pclause* W,
                                                const unsigned* table, unsigned lane) {
    const unsigned lt = lanemask_lt();
    ChoiceTracker trk;
    int out = 0;
    bool confl = false;
    for (int b = 0; b < nb; b += 32) {
        const int idx = b + (int)lane;
        const bool valid = idx < nb;
        pclause c = valid ? base[idx] : make_uint2(kTomb, kTomb);
        bool sat = false;
        if (kTable && valid) {
            unsigned f[3] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu};
            unsigned zc = (c.y >> 16) & 3u;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (f[j] != 0u) {
                    const unsigned val = table_get(table, f[j] >> 1);
                    const unsigned truth = (f[j] & 1u) ? 2u : 1u;  // variable value that makes the literal true
                    if (val == truth) sat = true;
                    else if (val != 0u) { f[j] = 0u; ++zc; }
                }
            }
            c.x = f[0] | (f[1] << 16);
            c.y = f[2] | (zc << 16);
        }
        const bool emit = valid && !sat;
        confl |= emit && ((c.y >> 16) & 3u) == 3u;
        const unsigned emask = __ballot_sync(kFullMask, emit);
        if (trk.lit2 == 0) trk.update(emit, c.x, c.y);
        if (emit) W[out + __popc(emask & lt)] = c;
        out += __popc(emask);
    }
    __syncwarp();
    pad_tail(W, out, lane);   // the streaming passes rely on a tombstone-padded tail
    SetInfo s;
    s.n = out;
    s.conflict = __any_sync(kFullMask, confl) ? 1 : 0;
    trk.finish(s);
    return s;
}

}  // namespace ndp
