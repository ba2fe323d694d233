// ndp_index_group.cuh -- the index DFS with SEVERAL searches per warp ("lane groups").
//
// dfs_index_kernel (ndp_index.cuh) runs one piece per warp.  99.6 % of its nodes are unit propagations whose
// only lane-parallel work is a row of ~4 word-merged occurrence entries; everything else -- finding the next
// unit clause, naming its literal, the trail, the counters -- is the same ~150 instructions whichever lane
// executes them, and ncu shows the kernel bound by instruction issue (profiles/r2_dfs_index_kernel_v7_*).
// Here a warp is cut into G = 32 / L groups of L lanes (L = 8 or 16) and each group searches its own piece in
// its own shared-memory state slot.  The groups walk in LOCKSTEP through the unit-propagation step, so those
// ~150 issue slots advance G nodes instead of one.  Everything that is not a unit step -- fetching a piece,
// a decision (classify both children, snapshot), a backtrack (restore, re-derive), finishing a piece -- is a
// SERVICE: the whole warp serves one group at a time with the 32-lane primitives of ndp_index.cuh on that
// group's slot, exactly the code path of the one-piece kernel, while the other groups wait (services are
// 0.7 % of the nodes).  Same search, same order, same counters: this is a scheduling change only, and the
// parity suite runs every DFS test on it.
//
// Per-group scalars (live count, literal, trail length, counters ...) are held redundantly by every lane of
// the group, so group-uniform control flow needs no broadcast; a service reads them from the group's first
// lane with one shuffle each and writes results back under `grp == g`.
#pragma once
#include "ndp_index.cuh"

namespace ndp {

enum : int { GS_FETCH = 0, GS_UNIT = 1, GS_CHOOSE = 2, GS_DECIDE = 3, GS_BACK = 4, GS_FINISH = 5, GS_DONE = 6 };

// kGlobal: the searchers' state slots live in global memory (p.gstate, served by L1/L2) instead of shared
// memory: the number of pieces in flight is then bounded by warps x groups, not by 228 KB / slot size.
template <int L, bool kGlobal = false, int kMinBlocks = 8>
__global__ void __launch_bounds__(128, kMinBlocks) dfs_index_group_kernel(const __grid_constant__ IndexParams p) {
    constexpr int G = 32 / L;
    constexpr unsigned kGroupBits = L == 32 ? 0xffffffffu : ((1u << L) - 1u);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const IndexRoot& ix = p.ix;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned grp = lane / L, gl = lane % L, gbase = grp * L;
    const unsigned gmask = kGroupBits << gbase;
    const unsigned searcher0 = (blockIdx.x * (blockDim.x >> 5) + warp) * G;   // this warp's first searcher
    unsigned char* warp_smem = kGlobal ? p.gstate + (size_t)searcher0 * p.warp_bytes : smem_raw + (size_t)warp * G * p.warp_bytes;
    // this lane's own group
    unsigned* B = reinterpret_cast<unsigned*>(warp_smem + (size_t)grp * p.warp_bytes);
    unsigned* table = B + ix.table_off;
    const unsigned tbase = (searcher0 + grp) * (unsigned)p.trail_cap;

    // group-uniform state
    int status = GS_FETCH, n_live = 0, lit = 0, ntrail = 0, verdict = 0;
    int npend = 0, snap_low = 0, snap_n = 0;      // snap_n: lane k of the group holds |A| of snapshot k
    unsigned long long nodes = 0, evals = 0, gq = 0;
    unsigned rebuilds = 0, q = 0, o = 0;
    unsigned steps = 0;

    for (;;) {
        // ------------------------------------------------------------------ services, one group at a time
        unsigned need = __ballot_sync(kFullMask, status != GS_UNIT && status != GS_DONE);
        while (need) {
            const int g = (__ffs(need) - 1) / L;
            const int src = g * L;
            const int st = __shfl_sync(kFullMask, status, src);
            const bool mine = (int)grp == g;
            unsigned* Bq = reinterpret_cast<unsigned*>(warp_smem + (size_t)g * p.warp_bytes);
            unsigned* tableq = Bq + ix.table_off;
            unsigned* pendq = reinterpret_cast<unsigned*>(warp_smem + (size_t)g * p.warp_bytes + p.off_pend);
            const unsigned tbq = (searcher0 + (unsigned)g) * (unsigned)p.trail_cap;
            unsigned* snapq = p.snaps + (size_t)(searcher0 + (unsigned)g) * p.snap_depth * ix.slot_words;
            if (st == GS_FETCH) {
                unsigned qq = 0;
                if (lane == 0) {
                    qq = atomicAdd(p.next, 1u);
                    if (p.progress) *(volatile unsigned*)p.progress = qq;
                }
                qq = __shfl_sync(kFullMask, qq, 0);
                if (qq >= p.n_local || (p.max_steps && steps >= p.max_steps)) {
                    if (mine) status = GS_DONE;
                } else {
                    const unsigned oo = p.orig ? (unsigned)p.orig[qq] : qq;
                    const unsigned long long gg = p.first + (unsigned long long)oo * p.stride;
                    int skip = 0;
                    if (lane == 0) skip = (p.early_exit && *(volatile unsigned long long*)p.best < gg) ||
                                          *(volatile unsigned int*)(p.best_piece + oo) < qq;
                    if (__shfl_sync(kFullMask, skip, 0)) {
                        if (lane == 0) { p.verdict[qq] = 2; p.nodes[qq] = 0; p.evals[qq] = 0; p.rebuilds[qq] = 0; }
                        // stays in GS_FETCH: the next round of the service loop pulls another unit
                    } else {
                        const BfsNode nd = p.meta[qq];
                        __syncwarp();
                        if (nd.kind != K_LEAF) {
                            slot_copy(Bq, p.slot_ptr[qq], ix.slot_words, lane);
                            for (int w = lane; w < p.pend_words; w += 32) pendq[w] = 0u;
                        }
                        __syncwarp();
                        if (mine) {
                            q = qq; o = oo; gq = gg;
                            n_live = nd.n; lit = nd.lit;
                            ntrail = 0; nodes = 0; evals = 0; rebuilds = 0; npend = 0; snap_low = 0; verdict = 0;
                            if (nd.kind == K_LEAF) { verdict = 1; status = GS_FINISH; }                     // the split reached the model
                            else if (nd.n == 0 || nd.lit == 0) { nodes = 1; verdict = 1; status = GS_FINISH; }   // choice()==0 (NDP:817-823)
                            else status = nd.kind == K_UNIT ? GS_UNIT : GS_DECIDE;
                        }
                    }
                }
            } else if (st == GS_CHOOSE || st == GS_DECIDE) {
                // a decision node: lit = the variable (GS_CHOOSE: still to be named by choice())
                const int nl = __shfl_sync(kFullMask, n_live, src);
                int var = __shfl_sync(kFullMask, lit, src);
                const int nt = __shfl_sync(kFullMask, ntrail, src);
                const int np_ = __shfl_sync(kFullMask, npend, src);
                if (st == GS_CHOOSE) var = index_choice_decide(ix, Bq, tableq, lane);
                bool la_conf, ra_conf;
                int la_n, ra_n;
                index_classify(ix, Bq, nl, var, la_n, la_conf, ra_n, ra_conf, lane);
                if (la_n == 0 || ra_n == 0) {                       // an empty child is a model; LA wins first (NDP:843-850, 863-870)
                    if (lane == 0) p.trails[tbq + (unsigned)nt] = (short)(la_n == 0 ? var : -var);
                    if (mine) { ++nodes; evals += (unsigned long long)nl; ++ntrail; verdict = 1; status = GS_FINISH; }
                } else {
                    const int t = !ra_conf ? -var : (!la_conf ? var : 0);   // RA is explored first
                    if (t == 0) {
                        if (mine) { ++nodes; evals += (unsigned long long)nl; status = GS_BACK; }   // both children conflict
                    } else {
                        if (!ra_conf && !la_conf) {                 // both open: LA stays pending, keep the state before the branch
                            if (p.snap_depth > 0) {
                                const int k = np_ % p.snap_depth;
                                __syncwarp();
                                slot_copy(snapq + (size_t)k * ix.slot_words, Bq, ix.slot_words, lane);
                                if ((int)lane == src + k) snap_n = nl;
                                if (mine) snap_low = max(min(snap_low, np_), np_ - p.snap_depth + 1);
                            }
                            if (mine) ++npend;
                            if (lane == 0) pendq[nt >> 5] |= 1u << (nt & 31);
                        }
                        if (mine) { lit = t; status = GS_UNIT; }    // the unit step below counts and applies the node
                    }
                }
                __syncwarp();
            } else if (st == GS_BACK) {
                // pop to the deepest decision whose LA sibling is still pending
                const int nt = __shfl_sync(kFullMask, ntrail, src);
                const int np_ = __shfl_sync(kFullMask, npend, src);
                const int sl = __shfl_sync(kFullMask, snap_low, src);
                const unsigned qq = __shfl_sync(kFullMask, q, src);
                __syncwarp();
                int pos = -1;
                for (int w = (nt - 1) >> 5; w >= 0 && nt > 0; --w) {
                    unsigned bits = pendq[w];
                    if (w == ((nt - 1) >> 5)) {
                        const int top = (nt - 1) & 31;
                        if (top != 31) bits &= (2u << top) - 1u;
                    }
                    if (bits) { pos = w * 32 + 31 - __clz(bits); break; }
                }
                if (pos < 0) {
                    if (mine) { verdict = 0; status = GS_FINISH; }   // stack empty: UNSAT
                } else {
                    __syncwarp();
                    if (lane == 0) {
                        pendq[pos >> 5] &= ~(1u << (pos & 31));
                        p.trails[tbq + (unsigned)pos] = (short)(-p.trails[tbq + (unsigned)pos]);   // -i -> +i: now the LA sibling
                    }
                    const int np2 = np_ - 1;
                    int kind2, lit2, nl2;
                    int sl2 = sl;
                    if (p.snap_depth > 0 && np2 >= sl) {
                        const int k = np2 % p.snap_depth;
                        __syncwarp();
                        __threadfence_block();
                        slot_copy(Bq, snapq + (size_t)k * ix.slot_words, ix.slot_words, lane);
                        const int la = (int)p.trails[tbq + (unsigned)pos];
                        const int n_before = __shfl_sync(kFullMask, snap_n, src + k);
                        __syncwarp();
                        if (lane == 0) table_set(tableq, la);
                        bool conflict;
                        nl2 = index_assign<true, true>(ix, Bq, n_before, la, conflict, lane);   // open and non-empty: it was left pending
                        index_choice<true>(ix, Bq, tableq, nl2, lane, kind2, lit2);
                    } else {
                        sl2 = min(sl, np2);   // siblings pushed from here on have intact snapshots again
                        const unsigned* slot = p.slot_ptr[qq];
                        for (int w = lane; w < ix.table_words; w += 32) tableq[w] = __ldg(&slot[ix.table_off + w]);
                        __syncwarp();
                        __threadfence_block();
                        for (int j = lane; j < pos + 1; j += 32) {
                            const int l = p.trails[tbq + (unsigned)j];
                            const unsigned v = (unsigned)(l < 0 ? -l : l);
                            atomicOr(&tableq[v >> 4], (l > 0 ? 1u : 2u) << ((v & 15u) * 2u));
                        }
                        __syncwarp();
                        nl2 = index_rebuild(ix, Bq, tableq, lane);
                        index_choice<true>(ix, Bq, tableq, nl2, lane, kind2, lit2);
                    }
                    if (mine) {
                        ntrail = pos + 1; npend = np2; ++rebuilds; snap_low = sl2;
                        n_live = nl2; lit = lit2;
                        status = kind2 == K_UNIT ? GS_UNIT : GS_DECIDE;
                    }
                }
                __syncwarp();
            } else {   // GS_FINISH: record the piece; a model is parked if it is the lowest so far
                const int vd = __shfl_sync(kFullMask, verdict, src);
                const int nt = __shfl_sync(kFullMask, ntrail, src);
                const unsigned qq = __shfl_sync(kFullMask, q, src);
                const unsigned oo = __shfl_sync(kFullMask, o, src);
                if (vd == 1) {
                    int park = 0;
                    if ((int)lane == src) {
                        const unsigned old_piece = atomicMin(p.best_piece + oo, qq);
                        const unsigned long long old = atomicMin(p.best, gq);
                        if (gq < old) {
#pragma unroll
                            for (int d = 0; d < 8; ++d)   // constant indices: the parameter block stays in constant memory
                                if (d < p.n_peers) atomicMin_system(p.peer_best[d], gq);
                        }
                        park = qq < old_piece && gq <= old;
                    }
                    park = __shfl_sync(kFullMask, park, src);
                    if (park) {   // so far the first model of the lowest SAT subproblem: park it in this searcher's slot
                        __syncwarp();
                        __threadfence_block();
                        int32_t* out = p.models + (size_t)(searcher0 + (unsigned)g) * p.trail_cap;
                        for (int j = lane; j < nt; j += 32) out[j] = (int32_t)p.trails[tbq + (unsigned)j];
                        if (lane == 0) { p.model_sub[searcher0 + (unsigned)g] = qq; p.model_len[searcher0 + (unsigned)g] = nt; }
                    }
                }
                if ((int)lane == src) {
                    p.verdict[qq] = (uint8_t)vd;
                    p.nodes[qq] = nodes;
                    p.evals[qq] = evals;
                    p.rebuilds[qq] = rebuilds;
                }
                if (mine) status = GS_FETCH;
                __syncwarp();
            }
            need = __ballot_sync(kFullMask, status != GS_UNIT && status != GS_DONE);
        }
        if (__all_sync(kFullMask, status == GS_DONE)) break;

        // ------------------------------------------------------------------ one unit step, every group in lockstep
        // (a group that is GS_DONE idles through it with everything predicated off)
        if ((++steps & 63u) == 0u) {     // every 64 steps: has a lower model become known?  (one vote: the group must agree)
            const bool gone = status == GS_UNIT && ((p.early_exit && *(volatile unsigned long long*)p.best < gq) ||
                                                    *(volatile unsigned int*)(p.best_piece + o) < q);
            if ((__ballot_sync(kFullMask, gone) & gmask) || (p.max_steps && steps >= p.max_steps && status == GS_UNIT)) {
                verdict = 2;             // abandon the piece
                status = GS_FINISH;
            }
        }
        const bool run = status == GS_UNIT;
        const int t = lit;
        const unsigned v = (unsigned)(t < 0 ? -t : t);
        uint2 hdr = make_uint2(0u, 0u);
        if (run) {
            ++nodes;                                   // one Satisfy_iterative loop iteration (NDP:801-806)
            evals += (unsigned long long)n_live;       // one ResolutionStepWithConflict (NDP:831)
            p.trails[tbase + (unsigned)ntrail] = (short)t;   // every lane of the group stores the same value
            table_set(table, t);
            ++ntrail;
            hdr = __ldg(&ix.row_hdr[v]);
        }
        const unsigned cnt = hdr.y & 0x7fffffffu;
        unsigned acc = 0;
        if (gl < cnt) acc = index_assign_entry<true, true>(ix, B, __ldg(&ix.rows[hdr.x + gl]), t > 0);
        if (__any_sync(kFullMask, cnt > (unsigned)L)) {        // a row longer than the group: more rounds (rare)
            const unsigned maxcnt = __reduce_max_sync(kFullMask, cnt);
#pragma unroll 1
            for (unsigned base = L; base < maxcnt; base += L) {
                const unsigned i = base + gl;
                if (i < cnt) {
                    const unsigned a2 = index_assign_entry<true, true>(ix, B, __ldg(&ix.rows[hdr.x + i]), t > 0);
                    acc = (acc + (a2 & (kConflBit - 1u))) | (a2 & kConflBit);
                }
            }
        }
        __syncwarp();
        // group-wide sum by butterfly: redux.sync with four different member masks compiles to a loop over the masks
        // (ncu: 4 REDUX per step); log2(L) shuffles do it in one pass (flags: <= L x 2^26 stays below 2^32)
        unsigned total = acc;
#pragma unroll
        for (int off = 1; off < L; off <<= 1) total += __shfl_xor_sync(kFullMask, total, off);
        if (run) {
            n_live -= (int)(total & (kConflBit - 1u));
            if (n_live == 0) { verdict = 1; status = GS_FINISH; }        // empty branch => model (NDP:843-852 / 863-872)
            else if ((total >> 26) != 0u) status = GS_BACK;              // a clause reached three zeros
        }
        // ---- choice(): the first unit clause through the B[2] summary, L summary words per round
        // Each lane reads up to kRounds summary words (gl, gl + L, ...: independent loads), one ballot per round names
        // the group's first non-zero word; no loop-carried vote, no second trip through the planes.
        const bool look = status == GS_UNIT;
        int c = -1;
        {
            constexpr int kRounds = 32 / L;              // sum_words <= 32
            unsigned sw_sel = 0u;                        // the summary word of this lane in the round that hit
            int hit = -1;                                // index (in words of the summary) of the first non-zero word
#pragma unroll
            for (int r = 0; r < kRounds; ++r) {
                if (r * L >= ix.sum_words) break;        // warp-uniform
                const int i = r * L + (int)gl;
                const unsigned sw = (look && i < ix.sum_words) ? B[ix.sum_off + i] : 0u;
                const unsigned m = (__ballot_sync(kFullMask, sw != 0u) >> gbase) & kGroupBits;
                if (hit < 0 && m) { hit = r * L + __ffs(m) - 1; sw_sel = sw; }
            }
            const unsigned word = __shfl_sync(kFullMask, sw_sel, gbase + (hit >= 0 ? (unsigned)hit % L : 0u));
            if (look && hit >= 0) {
                const int w = hit * 32 + __ffs(word) - 1;
                c = w * 32 + __ffs(B[w] & B[ix.bm_words + w]) - 1;
            }
        }
        // ---- the literal of that clause: the LAST one still there; lanes 0..2 of the group test one each
        {
            const bool have = status == GS_UNIT && c >= 0;
            const pclause cl = __ldg(&ix.root[have ? c : 0]);
            unsigned f = ((gl >= 2u ? cl.y : cl.x) >> ((gl & 1u) << 4)) & 0xffffu;
            if (gl > 2u || !have) f = 0u;
            const bool there = f != 0u && ((table[f >> 5] >> (f & 30u)) & 3u) == 0u;
            const unsigned m = (__ballot_sync(kFullMask, there) >> gbase) & kGroupBits;
            const unsigned code = __shfl_sync(kFullMask, f, gbase + (m ? 31 - __clz((int)m) : 0));
            if (status == GS_UNIT) {
                if (have) lit = code_lit(code);
                else status = GS_CHOOSE;                // no unit clause: a decision (service)
            }
        }
    }
}

}  // namespace ndp
