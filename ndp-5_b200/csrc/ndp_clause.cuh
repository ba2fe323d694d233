// ndp_clause.cuh -- packed clause representation and the warp-wide resolution pass.
//
// A clause is the reference's Clause3 (ClauseSetPool.hpp:37-39: three int literals, 0 = "already
// false / padding") packed into 8 bytes: three int16 fields + one zero pad, so that one lane
// moves one clause with a single 64-bit access and a warp moves 32 clauses (256 B) per
// instruction.  Clause sets are dense arrays of these records ("slots" of a device arena).
//
// warp_apply() is the device restatement of ResolutionStepWithConflict for ONE branch
// (NDP-5_6_7.cpp:654-697): make literal t true => clauses holding t are satisfied and dropped,
// occurrences of -t become 0, emission order is preserved (stable warp-scan compaction), and
// an emitted {0,0,0} raises the branch's conflict flag.  The same pass also produces what the
// next call of choice() (NDP-5_6_7.cpp:753-781) would return for the emitted set, by
// tracking the first emitted clause with two zeros, the first with one zero, and the first
// clause at all -- so the branching variable of the child is known without another scan.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ndp {

typedef uint2 pclause;  // x = l0 | l1 << 16 ; y = l2 (upper half zero)

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kMaxLit = 32767;

enum : int { K_EMPTY = 0, K_UNIT = 1, K_DECIDE = 2 };

// What choice() yields for a clause set, plus how it got there:
//  K_EMPTY : the set is empty (choice() == 0, NDP:780)
//  K_UNIT  : a clause with two zeros exists; lit = its (signed) remaining literal (NDP:756-766)
//  K_DECIDE: lit = the variable (>= 0) from the first one-zero clause's last literal
//            (NDP:767-777) or |A[0].l[0]| (NDP:778-779); lit == 0 only when A[0] is {0,0,0}.
struct SetInfo {
    int n;         // clauses in the set
    int conflict;  // some clause is {0,0,0}
    int kind;
    int lit;
};

__host__ __device__ __forceinline__ pclause pack_clause(int l0, int l1, int l2) {
    pclause c;
    c.x = ((unsigned)l0 & 0xffffu) | (((unsigned)l1 & 0xffffu) << 16);
    c.y = ((unsigned)l2 & 0xffffu);
    return c;
}
__host__ __device__ __forceinline__ int sext16(unsigned f) { return (int)(short)(unsigned short)f; }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Running answer to "what would choice() return for the clauses emitted so far".
struct ChoiceTracker {
    int lit2 = 0;        // literal of the first emitted two-zero clause (0 = none yet)
    int lit1 = 0;        // last literal of the first emitted one-zero clause (0 = none yet)
    int first = 0;       // l[0] of the first emitted clause
    bool have_first = false;

    // f0,f1,f2: this lane's post-resolution fields; emit: lane emits; zc: its zero count.
    __device__ __forceinline__ void update(unsigned emask, bool emit, int zc, unsigned f0, unsigned f1, unsigned f2) {
        if (!have_first && emask) {
            first = sext16(__shfl_sync(kFullMask, f0, __ffs(emask) - 1));
            have_first = true;
        }
        if (lit2 == 0) {
            unsigned m2 = __ballot_sync(kFullMask, emit && zc == 2);
            if (m2) {
                lit2 = sext16(__shfl_sync(kFullMask, f0 | f1 | f2, __ffs(m2) - 1));
            } else if (lit1 == 0) {
                unsigned m1 = __ballot_sync(kFullMask, emit && zc == 1);
                if (m1) lit1 = sext16(__shfl_sync(kFullMask, f2 ? f2 : f1, __ffs(m1) - 1));
            }
        }
    }
    __device__ __forceinline__ void finish(SetInfo& s) const {
        if (s.n == 0) { s.kind = K_EMPTY; s.lit = 0; }
        else if (lit2) { s.kind = K_UNIT; s.lit = lit2; }
        else { s.kind = K_DECIDE; int v = lit1 ? lit1 : first; s.lit = v < 0 ? -v : v; }
    }
};

// One branch of the resolution step over src[0..n): literal t becomes true.
//  kWrite: emit the surviving clauses to dst (dst may equal src: the compaction is stable and
//          only ever writes at or below the position it has already read, so it is safe in
//          place); without kWrite the pass only classifies the branch.
// All 32 lanes must call it with identical arguments; the result is warp-uniform.
template <bool kWrite, int kUnroll = 4>
__device__ __forceinline__ SetInfo warp_apply(const pclause* src, int n, pclause* dst, int t, unsigned lane) {
    const unsigned tu = (unsigned)t & 0xffffu;
    const unsigned ntu = (unsigned)(-t) & 0xffffu;
    const unsigned lt = lanemask_lt();
    ChoiceTracker trk;
    int out = 0;
    bool confl = false;
    for (int base = 0; base < n; base += 32 * kUnroll) {
        pclause c[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int idx = base + u * 32 + (int)lane;
            c[u] = idx < n ? src[idx] : make_uint2(tu, 0u);  // padding lanes look satisfied
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (base + u * 32 < n) {  // warp-uniform
                unsigned f0 = c[u].x & 0xffffu, f1 = c[u].x >> 16, f2 = c[u].y & 0xffffu;
                bool sat = (f0 == tu) | (f1 == tu) | (f2 == tu);
                f0 = f0 == ntu ? 0u : f0;
                f1 = f1 == ntu ? 0u : f1;
                f2 = f2 == ntu ? 0u : f2;
                int zc = (f0 == 0u) + (f1 == 0u) + (f2 == 0u);
                bool emit = !sat;
                confl |= emit && zc == 3;
                unsigned emask = __ballot_sync(kFullMask, emit);
                trk.update(emask, emit, zc, f0, f1, f2);
                if (kWrite && emit) dst[out + __popc(emask & lt)] = make_uint2(f0 | (f1 << 16), f2);
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

// 2-bit-per-variable assignment table: 0 unassigned, 1 true, 2 false.
__device__ __forceinline__ unsigned table_get(const unsigned* table, unsigned var) {
    return (table[var >> 4] >> ((var & 15u) * 2u)) & 3u;
}

// Re-materialise a clause set from its base under the assignment in `table`
// (the stable-filter equivalence of SURVEY.md §8a: a node's set == its base filtered by the
// assignment).  With kTable == false it is a plain copy that also computes the SetInfo.
template <bool kTable>
__device__ __forceinline__ SetInfo warp_rebuild(const pclause* __restrict__ base, int nb, pclause* W,
                                                const unsigned* table, unsigned lane) {
    const unsigned lt = lanemask_lt();
    ChoiceTracker trk;
    int out = 0;
    bool confl = false;
    for (int b = 0; b < nb; b += 32) {
        int idx = b + (int)lane;
        bool valid = idx < nb;
        pclause c = valid ? base[idx] : make_uint2(0u, 0u);
        unsigned f[3] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu};
        bool sat = false;
        if (kTable) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                int l = sext16(f[j]);
                if (l != 0) {
                    unsigned val = table_get(table, (unsigned)(l < 0 ? -l : l));
                    unsigned truth = l > 0 ? 1u : 2u;  // value of the variable that makes l true
                    if (val == truth) sat = true;
                    else if (val != 0u) f[j] = 0u;
                }
            }
        }
        int zc = (f[0] == 0u) + (f[1] == 0u) + (f[2] == 0u);
        bool emit = valid && !sat;
        confl |= emit && zc == 3;
        unsigned emask = __ballot_sync(kFullMask, emit);
        trk.update(emask, emit, zc, f[0], f[1], f[2]);
        if (emit) W[out + __popc(emask & lt)] = make_uint2(f[0] | (f[1] << 16), f[2]);
        out += __popc(emask);
    }
    SetInfo s;
    s.n = out;
    s.conflict = __any_sync(kFullMask, confl) ? 1 : 0;
    trk.finish(s);
    return s;
}

}  // namespace ndp
