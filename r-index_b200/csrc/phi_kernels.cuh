// phi_kernels.cuh — sm_100a Phi expansion: the locate_all loop of the reference
// (internal/r_index.hpp:340-351: OCC = SA[hi], Phi(SA[hi]), Phi^2(SA[hi]), ...) for a whole batch.
//
// TWO PASSES when the index carries a seed table (Phi^SEG, flat_layout.hpp: JumpTable), one otherwise:
//
//   phi_expand_kernel<.., SEEDED=true>   "seed pass", one lane per chain: writes the toehold, walks the chain's
//        head up to the next 128-byte line of the OUTPUT array (<= 16 occurrences), then hops along the chain
//        SEG occurrences at a time (one seed-table lookup per hop) and appends one 16-byte ITEM per hop to a
//        list: (first output slot, number of occurrences that follow, seed = the occurrence on that slot).
//   phi_window_kernel                    "window pass", one lane per item at a time (persistent lanes): from the
//        seed, every lookup yields D occurrences that leave as one aligned 32-byte sector.
// Both kernels are PERSISTENT (grid sized to the machine) and read the batch totals from device memory, so a
// locate call is queued without a host round trip between the search and the expansion.
//
// The single-pass form's time is (longest chain) x (load latency) with half-empty warps (chain lengths
// differ by orders of magnitude inside a warp); the two-pass form turns the batch into uniform
// 16-lookup work items, 1.6 M of them on config C2 instead of 143 k chains.
//
// One LANE per chain. A chain is the part of a pattern's SA range that lies inside one BWT run
// (ranges are cut at run boundaries: every run end is a free toehold, SA = samples_last[run]+1, the
// identity behind r_index.hpp:489,533), so a batch exposes (#patterns x #runs overlapped) independent
// chains instead of #patterns.
//
// One record LOOKUP yields D occurrences: the PhiTable (flat_layout.hpp) stores, per piece of the
// refined translation, the deltas of Phi^1..Phi^D, so from SA[x] one bucket record gives
// SA[x-1..x-D]; they leave as ONE aligned vector store (256-bit for D = 4).
//
// The walk is a per-lane STATE MACHINE with exactly one RW-word load per loop iteration, whatever the
// lane is doing (reading the bucket record of its current position, or probing a piece entry while it
// searches a crowded bucket): a lane's slow path costs that lane extra iterations, not the warp.
// All value arithmetic is done in the table's word type (u32 when n < 2^32-1): the kernel was
// measured issue-bound at 113 instructions per iteration with 64-bit arithmetic.
#pragma once
#include <type_traits>
#include "search_kernels.cuh"

namespace rigk {

// Timing-only diagnostics of the window pass (profiles/r2_window_bound.txt): compiled in with -DRIG_WINDOW_DIAG, then
// selected by RIG_VARIANT bits 15-17 (stores redirected into one L2-resident window / no stores / no dependent
// lookups; all three give WRONG output). The product build folds them away.
#ifdef RIG_WINDOW_DIAG
#define RIG_DIAG(ix) ((ix).pad)
#else
#define RIG_DIAG(ix) 0u
#endif

template <bool KEEP>
__device__ __forceinline__ void ldg256(const void* p, u64& a, u64& b, u64& c, u64& d) {
    if (KEEP)  // ask L2 to evict these lines last: the streamed occurrence output competes for the same sets
        asm volatile("ld.global.nc.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    else
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void stg256_stream(void* p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.u64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// Load the RW-word entry at `p` (entry-size aligned). packed (64-bit words, RW = 8 only): the entry is 32 bytes,
// 4 x 40-bit deltas | 40-bit s1/start | 32-bit nxt | 24-bit cnt, unpacked into the same word positions.
template <typename WT, int RW, bool KEEP>
__device__ __forceinline__ void load_entry(const void* p, WT (&w)[RW], bool packed = false) {
    if constexpr (sizeof(WT) == 8 && RW == 8) {
        if (packed) {
            u64 q0, q1, q2, q3;
            ldg256<KEEP>(p, q0, q1, q2, q3);
            const u64 m40 = 0xFFFFFFFFFFull;
            w[0] = q0 & m40;
            w[1] = ((q0 >> 40) | (q1 << 24)) & m40;
            w[2] = (q1 >> 16) & m40;
            w[3] = ((q1 >> 56) | (q2 << 8)) & m40;
            w[4] = ((q2 >> 32) | (q3 << 32)) & m40;
            if (w[4] == m40) w[4] = ~0ull;  // the "no piece begins inside this bucket" sentinel
            w[5] = (q3 >> 8) & 0xFFFFFFFFull;
            w[6] = q3 >> 40;
            w[7] = 0;
            return;
        }
    }
    if constexpr (sizeof(WT) == 4) {
        if constexpr (RW == 4) {
            const uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
            w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
        } else {
#pragma unroll
            for (int k = 0; k < RW / 8; ++k) {
                u64 a, b, c, d;
                ldg256<KEEP>(reinterpret_cast<const u32*>(p) + 8 * k, a, b, c, d);
                w[8 * k + 0] = (u32)a; w[8 * k + 1] = (u32)(a >> 32); w[8 * k + 2] = (u32)b; w[8 * k + 3] = (u32)(b >> 32);
                w[8 * k + 4] = (u32)c; w[8 * k + 5] = (u32)(c >> 32); w[8 * k + 6] = (u32)d; w[8 * k + 7] = (u32)(d >> 32);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < RW / 4; ++k)
            ldg256<KEEP>(reinterpret_cast<const u64*>(p) + 4 * k, w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
}

// OT = output word: u64 (the reference's ulint) or u32 (rig_locate_batch32, n < 2^32: half the store sectors)
template <typename WT, int D, typename OT>
__device__ __forceinline__ void store_group(OT* p, const WT (&e)[D]) {  // p is D*sizeof(OT)-byte aligned
    if constexpr (sizeof(OT) == 8) {
        if constexpr (D == 1) __stcs(p, (OT)e[0]);
        else if constexpr (D == 2) __stcs(reinterpret_cast<ulonglong2*>(p), make_ulonglong2((u64)e[0], (u64)e[1]));
        else {
#pragma unroll
            for (int k = 0; k < D / 4; ++k)
                stg256_stream(p + 4 * k, (u64)e[4 * k], (u64)e[4 * k + 1], (u64)e[4 * k + 2], (u64)e[4 * k + 3]);
        }
    } else {
        if constexpr (D == 1) __stcs(p, (OT)e[0]);
        else if constexpr (D == 2) __stcs(reinterpret_cast<uint2*>(p), make_uint2((u32)e[0], (u32)e[1]));
        else {
#pragma unroll
            for (int k = 0; k < D / 4; ++k)
                __stcs(reinterpret_cast<uint4*>(p) + k, make_uint4((u32)e[4 * k], (u32)e[4 * k + 1], (u32)e[4 * k + 2], (u32)e[4 * k + 3]));
        }
    }
}

// nxt / cnt of a bucket record: words D+1 and D+2, or — D = 6, the six-delta 32-byte entry — one shared word
// (nxt in the low 24 bits, cnt in the high 8; flat_layout.hpp: PhiTable)
template <typename WT, int D, int RW>
__device__ __forceinline__ u32 rec_nxt(const WT (&e)[RW]) { return D == 6 ? ((u32)e[7] & 0xFFFFFFu) : (u32)e[D + 1]; }
template <typename WT, int D, int RW>
__device__ __forceinline__ u32 rec_cnt(const WT (&e)[RW]) { return D == 6 ? ((u32)e[7] >> 24) : (u32)e[D + 2]; }

// Pull the Phi tables into L2 with one streaming pass (evict_last) before the walk: a cold table costs
// every chain a DRAM round trip per first touch, and with 32 lanes in lockstep almost every iteration
// of every warp would contain one. ~40 MB at HBM speed is a few microseconds.
__global__ void __launch_bounds__(256) l2_warm_kernel(const char* base, u64 bytes) {
    const u64 line = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    if (line < bytes) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(base + line));
}

// Start of a batch call, ONE launch instead of two memset nodes and a warm-up kernel: zero the call's counters and
// the workspace of the search kernel's offset scan, and pull the backward-search structures into L2 (warm_bytes = 0:
// they are too large to stay there). The search is a chain of dependent loads: with 32 lanes in lockstep one cold
// line per warp-step costs the whole warp a DRAM round trip.
__global__ void __launch_bounds__(256) prep_kernel(u64* z0, u64 n0, u64* z1, u64 n1, const char* base, u64 warm_bytes) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x, T = (u64)gridDim.x * blockDim.x;
    for (u64 i = t; i < n0; i += T) z0[i] = 0;
    for (u64 i = t; i < n1; i += T) z1[i] = 0;
    for (u64 line = t * 128; line < warm_bytes; line += T * 128) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(base + line));
}

// ---- walking a chain with the Phi^1..Phi^D table ---------------------------------------------------
// From v = SA[x] (already written at o[-1]) produce `remaining` further occurrences SA[x-1], SA[x-2], ...
// at o[0], o[1], ...; returns the last value produced (v itself when remaining == 0).
//
// Per step, from the current value v = SA[x]: e[t] = Phi^(t+1)(v) = (v + delta_t) mod n, where the
// deltas are those of the piece holding v. For t = 0 this is r_index::Phi (r_index.hpp:195-221):
// strict circular predecessor over the sorted run-first samples (sparse_sd_vector.hpp:107-112,153-157)
// and (prev_sample + delta) % n (:219), folded into one delta per piece.
template <typename WT, int D, bool KEEP, typename OT>
__device__ __forceinline__ WT walk_chain(const FlatDev& ix, WT v, OT* o, WT remaining) {
    static_assert(D != 6 || sizeof(OT) == 8, "six per entry: 64-bit output only");
    constexpr int RW = (D == 1) ? 4 : ((D <= 6) ? 8 : 16);
    const u32 ESZ = ix.phi.esz;  // entry size in bytes (RW words, or 32 when packed)
    const bool PK = ix.phi.packed != 0;
    constexpr bool W32 = sizeof(WT) == 4;
    const WT n = (WT)ix.n;
    const char* rec = reinterpret_cast<const char*>(ix.phi.rec);
    const char* pent = reinterpret_cast<const char*>(ix.phi.pent);
    const u32 shift = ix.phi.shift;
    // slots up to the next D-slot boundary go out singly; afterwards every emit is one aligned group
    u32 take = D - (u32)((reinterpret_cast<unsigned long long>(o) / sizeof(OT)) % D);
    bool searching = false;   // false: next load = bucket record of v; true: next load = piece entry `probe`
    u32 slo = 0, shi = 0;     // search interval of piece indices, invariant start[slo] <= v
    // Software-pipelined: the entry for the CURRENT state is already in flight / in registers when an
    // iteration starts; the iteration first decides (emit or narrow the search), computes the next
    // address and ISSUES THE NEXT LOAD, and only then writes this iteration's occurrences — so the
    // loop-carried chain between two loads is ~10 ALU ops and the stores overlap the load latency
    // (measured: 55% of the stall samples sat on the first use of the loaded entry).
    u32 probe = 0;
    WT e[RW];
    if (remaining > 0) load_entry<WT, RW, KEEP>(rec + (u64)(v >> shift) * ESZ, e, PK);
    while (remaining > 0) {
        bool emit;
        if (!searching) {
            emit = v < e[D];              // no piece begins inside the bucket at or below v
            slo = rec_nxt<WT, D, RW>(e); shi = slo + rec_cnt<WT, D, RW>(e) - 1;
            searching = !emit;
        } else if (slo == shi) {
            emit = true;                  // the entry just loaded is the answer
        } else if (e[D] <= v) {
            slo = probe; emit = (slo == shi);
        } else {
            shi = probe - 1; emit = false;
        }
        WT x[D];
        u32 cnt = 0;
        WT vn = v;
        if (emit) {
            searching = false;
            slo = shi = 0;
#pragma unroll
            for (int t = 0; t < D; ++t) {
                x[t] = v + e[t];
                if ((W32 && x[t] < v) || x[t] >= n) x[t] -= n;  // (v + delta) mod n; 32-bit: detect the carry
            }
            cnt = (take == D && remaining >= (WT)D) ? (u32)D : (u32)min((u64)take, (u64)remaining);
            vn = x[D - 1];
#pragma unroll
            for (int t = 0; t < D - 1; ++t)
                if ((u32)(t + 1) == cnt) vn = x[t];
        }
        const WT rem_next = remaining - (WT)cnt;
        // ---- next load (critical path) ----
        probe = (slo < shi) ? ((slo + shi + 1) >> 1) : slo;
        WT e2[RW];
        if (rem_next > 0)
            load_entry<WT, RW, KEEP>(searching ? (pent + (u64)probe * ESZ) : (rec + (u64)(vn >> shift) * ESZ), e2, PK);
        // ---- this iteration's occurrences (off the critical path) ----
        if (emit) {
            if (D != 6 && cnt == (u32)D) {
                store_group<WT, D, OT>(o, x);
            } else {   // partial group, or D = 6 (a group of six is not sector-aligned: the head walk is short, singles do)
#pragma unroll
                for (int t = 0; t < D; ++t)
                    if ((u32)t < cnt) __stcs(o + t, (OT)x[t]);
                take = D;
            }
            o += cnt;
        }
        v = vn;
        remaining = rem_next;
#pragma unroll
        for (int t = 0; t < RW; ++t) e[t] = e2[t];
    }
    return v;
}

// One hop of SEG occurrences: Phi^SEG(v) through the seed table (flat_layout.hpp: JumpTable): a 16-word
// bucket record resolving up to 6 pieces that begin inside the bucket; binary search over 2-word piece
// entries only in more crowded buckets. The hops of a chain are dependent DRAM-latency loads (the seed
// table is far larger than L2), so one hop must be one load.
template <typename WT>
__device__ __forceinline__ WT seed_hop(const FlatDev& ix, WT v) {
    constexpr bool W32 = sizeof(WT) == 4;
    WT w[16];
    load_entry<WT, 16, false>(reinterpret_cast<const char*>(ix.seed.rec) + (u64)(v >> ix.seed.shift) * (16 * sizeof(WT)), w);
    WT d = w[0];
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (v >= w[1 + 2 * i]) d = w[2 + 2 * i];   // starts ascend; unused slots hold ~0 > v
    if ((u32)w[14] > 6u && v >= w[11]) {  // piece nxt+5 starts at s_5 <= v: last piece of [nxt+5, nxt+cnt) with start <= v
        const WT* pe = reinterpret_cast<const WT*>(ix.seed.pent);
        u32 lo = (u32)w[13] + 5, hi = (u32)w[13] + (u32)w[14] - 1;
        while (lo < hi) {
            const u32 mid = (lo + hi + 1) >> 1;
            if (__ldg(pe + 2 * (u64)mid + 1) <= v) lo = mid; else hi = mid - 1;
        }
        d = __ldg(pe + 2 * (u64)lo);
    }
    WT x = v + d;
    if ((W32 && x < v) || x >= (WT)ix.n) x -= (WT)ix.n;
    return x;
}

__device__ __forceinline__ void stg128_stream(void* p, u64 a, u64 b) {
    // (the .L2::evict_first qualifier is only accepted on 256-bit stores; .cs is the 128-bit streaming form)
    asm volatile("st.global.cs.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}


// Device-side view of a locate call's counters (rig_index::d_counters): the expansion kernels read the totals the
// search produced instead of taking them as launch parameters, so the host can queue the whole call without
// waiting for the scan (no idle gap between the search and the expansion).
#define RIG_CTR_TOTAL 2    // occurrences of the batch
#define RIG_CTR_CHAINS 3   // Phi chains of the batch
#define RIG_CTR_ITEMS 6    // items appended by the seed pass
// Expansion of a SHARD of a planned batch (rig_expand_shard_dev): the arrays are passed shifted to the shard's first
// pattern and keep their batch-wide values; a second counter block holds the shard's totals and these two bases (both
// 0 for a whole batch).
#define RIG_CTR_OCC_BASE 9   // output offset of the first pattern
#define RIG_CTR_CH_BASE 10   // chain offset of the first pattern
#define RIG_CTR_BLOCK 16     // words per counter block

// Both kernels return at once when the host will not want the output (total > capacity) or when the item list
// could overflow (the host re-launches with a larger list: it evaluates the same two conditions from the same
// totals). An upper bound of the items: a chain of L occurrences gives at most (L - 1) / SEG + 1.
__device__ __forceinline__ bool expansion_enabled(const u64* ctr, u64 cap, u64 items_cap, u32 seg_shift, bool seeded,
                                                  u64& total, u64& chains) {
    total = __ldcg(ctr + RIG_CTR_TOTAL);
    chains = __ldcg(ctr + RIG_CTR_CHAINS);
    if (total > cap) return false;
    if (seeded && (total >> seg_shift) + chains > items_cap) return false;
    return true;
}

// Work item w -> (pattern p, run j): BWT positions [max(lo,start[j]), min(hi,start[j+1]-1)], walked
// from the top down. Output slot of SA[x] is occ_off[p] + (hi - x): locate_all order (r_index.hpp:340-351).
// PERSISTENT: the grid is sized to the machine (rig_index: SMs x resident CTAs), every warp strides over the
// chains, and their number comes from device memory.
//
// SEEDED = false: the lane produces its whole chain.
// SEEDED = true : the lane produces the chain's head up to the next 128-byte line of the output array
//                 (<= 16 occurrences), then cuts the rest of the chain into ITEMS of SEG = 1 << seg_shift
//                 slots: per item one seed_hop and one 16-byte entry in items[]: (first slot << 8 |
//                 occurrences after the seed, SEED = the occurrence on the item's first slot). The seed
//                 travels in the entry, not through the output array: an 8-byte store into a random output
//                 line is a DRAM read-modify-write here and a random sector read in pass 2, and the seed pass
//                 was measured bound by exactly that random-access rate. The chain's K entries are
//                 reserved up front with ONE atomic per warp (shuffle scan of K), so the reservation's
//                 latency hides behind the head walk. phi_window_kernel fills in the items.
#define RIG_ITEM_MASK 0x0000FFFFFFFFFFFFull   // an item word without its epoch tag (fused expansion)

// The chains wb .. of one warp-stride loop: toehold, head walk, items. TAG != 0: the fused kernel's consumers poll
// the item list while it is being written; each 16-byte entry then carries the call's epoch in the top 16 bits of
// BOTH words (slots and positions stay below 2^48) and is written with a device-scope relaxed store, so a reader
// accepts an entry only when both tags are the current epoch.
template <typename WT, int D, bool KEEP, bool SEEDED, typename OT>
__device__ __forceinline__ void produce_chains(const FlatDev& ix, u64 N, const u64* __restrict__ ch_off,
                                               const u64* __restrict__ occ_off, const u64* __restrict__ lo_in,
                                               const u64* __restrict__ hi_in, const u64* __restrict__ toe_in,
                                               const u64* __restrict__ jl_in, OT* __restrict__ out, u64* __restrict__ ctr,
                                               u64* __restrict__ items, u32 seg_shift, u64 total_chains, u64 tag,
                                               u32 block, u32 nblocks) {   // this CTA's rank among the nblocks producing CTAs
    constexpr u64 LINE = 128 / sizeof(OT);   // output slots per 128-byte line
    const int lane = threadIdx.x & 31;
    const u64 stride = (u64)nblocks * blockDim.x;
    const u64 occ_base = __ldcg(ctr + RIG_CTR_OCC_BASE), ch_base = __ldcg(ctr + RIG_CTR_CH_BASE);
    for (u64 wb = (u64)block * blockDim.x + (threadIdx.x & ~31u); wb < total_chains; wb += stride) {  // warp-uniform
        const u64 w = wb + lane;
        const bool active = w < total_chains;
        u64 g0 = 0, glast = 0, v0 = 0;
        if (active) {
            u64 a = 0, b = N;  // largest p with ch_off[p] <= w
            while (b - a > 1) {
                const u64 mid = (a + b) >> 1;
                if (__ldg(ch_off + mid) - ch_base <= w) a = mid; else b = mid;
            }
            const u64 p = a;
            const u64 L = __ldg(lo_in + p), H = __ldg(hi_in + p);
            const u64 j = __ldg(jl_in + p) + (w - (__ldg(ch_off + p) - ch_base));
            const u64 sj = ld_pos<WT>(ix.start, j), ej = (u64)ld_pos<WT>(ix.start, j + 1) - 1;
            const u64 top = min(H, ej), bot = max(L, sj);
            if (top == H) v0 = __ldg(toe_in + p);  // toehold carried by the backward search (r_index.hpp:482-545)
            else { v0 = (u64)ld_pos<WT>(ix.samples_last, j) + 1; if (v0 >= ix.n) v0 -= ix.n; }  // run end: SA = sample + 1
            g0 = __ldg(occ_off + p) - occ_base + (H - top);  // slot of the chain's first occurrence
            glast = g0 + (top - bot);             // slot of its last (a chain never exceeds n)
            __stcs(out + g0, (OT)v0);
        }
        if (!SEEDED) {
            if (active) walk_chain<WT, D, KEEP, OT>(ix, (WT)v0, out + g0 + 1, (WT)(glast - g0));
        } else {
            const u64 SEG = 1ull << seg_shift;
            const u64 a1 = (g0 + LINE - 1) & ~(LINE - 1);  // first line-aligned slot at or after g0
            const u64 K = (active && a1 <= glast) ? ((glast - a1) >> seg_shift) + 1 : 0;  // items of this chain
            // reserve K entries of items[]: inclusive warp scan, one atomic by the last lane
            u64 incl = K;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 t = __shfl_up_sync(RIG_FULL, incl, d);
                if (lane >= d) incl += t;
            }
            u64 wbase = 0;
            if (lane == 31 && incl) wbase = atomicAdd(ctr + RIG_CTR_ITEMS, incl);
            wbase = __shfl_sync(RIG_FULL, wbase, 31);
            if (!active) continue;
            const u64 pre_last = min(a1, glast);
            WT v = walk_chain<WT, D, KEEP, OT>(ix, (WT)v0, out + g0 + 1, (WT)(pre_last - g0));
            if (K) {
                ulonglong2* it = reinterpret_cast<ulonglong2*>(items) + wbase + (incl - K);
                u64 s = a1;
                for (;;) {
                    const u64 w0 = (s << 8) | min(SEG - 1, glast - s), w1 = (u64)v;
                    if (tag) asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" :: "l"(it), "l"(w0 | tag), "l"(w1 | tag) : "memory");
                    else __stcs(it, make_ulonglong2(w0, w1));
                    ++it;
                    s += SEG;
                    if (s > glast) break;
                    v = seed_hop<WT>(ix, v);
                }
            }
        }
    }
}

template <typename WT, int D, bool KEEP, bool SEEDED, typename OT>
__global__ void __launch_bounds__(256)
phi_expand_kernel(const FlatDev ix, u64 N, const u64* __restrict__ ch_off, const u64* __restrict__ occ_off,
                  const u64* __restrict__ lo_in, const u64* __restrict__ hi_in, const u64* __restrict__ toe_in,
                  const u64* __restrict__ jl_in, OT* __restrict__ out, u64* __restrict__ ctr, u64 cap,
                  u64* __restrict__ items, u64 items_cap, u32 seg_shift) {
    u64 total, total_chains;
    if (!expansion_enabled(ctr, cap, items_cap, seg_shift, SEEDED, total, total_chains)) return;
    produce_chains<WT, D, KEEP, SEEDED, OT>(ix, N, ch_off, occ_off, lo_in, hi_in, toe_in, jl_in, out, ctr, items, seg_shift, total_chains, 0,
                                        blockIdx.x, gridDim.x);
}

// One item per lane, the warp in lockstep until its slowest lane is done: left = slots of the item still to be written
// (the seed's included, 0 = no item), o = its first slot (sector-aligned: items start on a 128-byte line), v = the seed.
// Direct 32-byte sector stores; per-lane state machine with ONE table load per lane per trip (the bucket record of
// the next value, or the next probe of a crowded bucket), issued before the trip's stores.
// D = 6 (on request): a lookup yields six occurrences but a sector holds four, so two values wait in registers every
// other lookup: lookup A stores [x0..x3] and keeps x4, x5; lookup B stores [x4, x5, y0, y1] and [y2..y5] — three
// sector stores per two lookups, the same store count per occurrence as D = 4 with two thirds of its lookups.
// (A leaner form of this loop — full groups only in the main loop, the next entry loaded straight into the registers
// of the one consumed, 124 M instead of 157 M warp instructions on config C2 — was measured no faster: 0.300 vs
// 0.285 ms. The pass is not issue-bound; profiles/r2_window_bound.txt.)
template <typename WT, int D, bool KEEP, typename OT>
__device__ __forceinline__ void window_items_direct(const FlatDev& ix, u32 left, OT* o, WT v) {
    static_assert(D != 6 || sizeof(OT) == 8, "six per entry: 64-bit output only");
    constexpr int RW = (D == 1) ? 4 : ((D <= 6) ? 8 : 16);
    const u32 ESZ = ix.phi.esz;
    const bool PK = ix.phi.packed != 0;
    constexpr bool W32 = sizeof(WT) == 4;
    const WT n = (WT)ix.n;
    const char* rec = reinterpret_cast<const char*>(ix.phi.rec);
    const char* pent = reinterpret_cast<const char*>(ix.phi.pent);
    const u32 shift = ix.phi.shift;
    bool searching = false;
    u32 slo = 0, shi = 0, probe = 0;
    u32 pc = 0;          // values waiting for their sector, destined for slots o, o + 1, ..: D = 6: 0 or 2; 32-bit output: 0 or 4
    WT p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    WT e[RW];
    if (left > 1) load_entry<WT, RW, KEEP>(rec + (u64)(v >> shift) * ESZ, e, PK);
    while (__any_sync(RIG_FULL, left > 1)) {
        bool emit = false;
        WT g[D];
        u32 cnt = 0;
        WT vn = v;
        if (left > 1) {
            if (!searching) {
                emit = v < e[D];
                slo = rec_nxt<WT, D, RW>(e); shi = slo + rec_cnt<WT, D, RW>(e) - 1;
                searching = !emit;
            } else if (slo == shi) {
                emit = true;
            } else if (e[D] <= v) {
                slo = probe; emit = (slo == shi);
            } else {
                shi = probe - 1; emit = false;
            }
            if (RIG_DIAG(ix) == 3) emit = true;   // diagnostic: no dependent lookups (WRONG output, timing only)
            if (emit) {
                searching = false;
                slo = shi = 0;
                g[0] = v;
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    WT x = v + e[t];
                    if ((W32 && x < v) || x >= n) x -= n;
                    if (t < D - 1) g[t + 1] = x; else vn = x;
                }
                cnt = min(left, (u32)D);
            }
        }
        const u32 left_next = left - cnt;
        probe = (slo < shi) ? ((slo + shi + 1) >> 1) : slo;
        WT e2[RW];
        if (RIG_DIAG(ix) == 3) {
#pragma unroll
            for (int t = 0; t < RW; ++t) e2[t] = e[t];
        } else if (left_next > 1)
            load_entry<WT, RW, KEEP>(searching ? (pent + (u64)probe * ESZ) : (rec + (u64)(vn >> shift) * ESZ), e2, PK);
        if (emit) {
            if constexpr (D == 6) {
                if (pc == 0) {
                    if (cnt >= 4) {
                        stg256_stream(o, (u64)g[0], (u64)g[1], (u64)g[2], (u64)g[3]);
                        o += 4;
                        if (cnt == 6) { p0 = g[4]; p1 = g[5]; pc = 2; }
                        else if (cnt == 5) { __stcs(o, (u64)g[4]); o += 1; }
                    } else {
#pragma unroll
                        for (int t = 0; t < 3; ++t)
                            if ((u32)t < cnt) __stcs(o + t, (u64)g[t]);
                        o += cnt;
                    }
                } else {   // slots o, o + 1 hold p0, p1
                    if (cnt >= 2) {
                        stg256_stream(o, (u64)p0, (u64)p1, (u64)g[0], (u64)g[1]);
                        o += 4; pc = 0;
                        if (cnt == 6) {
                            stg256_stream(o, (u64)g[2], (u64)g[3], (u64)g[4], (u64)g[5]);
                            o += 4;
                        } else {
#pragma unroll
                            for (int t = 2; t < 5; ++t)
                                if ((u32)t < cnt) __stcs(o + (t - 2), (u64)g[t]);
                            o += cnt - 2;
                        }
                    } else {
                        __stcs(o, (u64)p0); __stcs(o + 1, (u64)p1);
                        if (cnt == 1) __stcs(o + 2, (u64)g[0]);
                        o += 2 + cnt; pc = 0;
                    }
                }
            } else if constexpr (sizeof(OT) == 4 && D == 4) {
                // 32-bit output: a sector holds eight positions = two lookups. The first group of a sector waits in
                // registers, the second completes it: one 32-byte store per two lookups (16-byte stores are half-sector
                // writes: measured 0.49 ms against 0.29 ms for the 64-bit output). Windows start on a 128-byte line, so
                // the groups of a lane alternate first / second.
                if (cnt == 4u) {
                    if (pc == 0) { p0 = g[0]; p1 = g[1]; p2 = g[2]; p3 = g[3]; pc = 4; }
                    else {
                        stg256_stream(o, (u64)p0 | ((u64)p1 << 32), (u64)p2 | ((u64)p3 << 32), (u64)g[0] | ((u64)g[1] << 32),
                                      (u64)g[2] | ((u64)g[3] << 32));
                        o += 8; pc = 0;
                    }
                } else {
                    if (pc) { __stcs(reinterpret_cast<uint4*>(o), make_uint4((u32)p0, (u32)p1, (u32)p2, (u32)p3)); o += 4; pc = 0; }
#pragma unroll
                    for (int t = 0; t < D - 1; ++t)
                        if ((u32)t < cnt) __stcs(o + t, (OT)g[t]);
                    o += cnt;
                }
            } else {
                if (cnt == (u32)D) {
                    if (RIG_DIAG(ix) == 0 || RIG_DIAG(ix) == 3) store_group<WT, D, OT>(o, g);
                    else if (RIG_DIAG(ix) == 1)   // diagnostic (RIG_VARIANT bit 15): every store lands in one 32 MB window — WRONG output, timing only
                        store_group<WT, D, OT>(reinterpret_cast<OT*>(ix.dbg) + ((reinterpret_cast<unsigned long long>(o) / sizeof(OT)) & 0x3FFFFCull), g);
                    else if (g[0] == (WT)0xFFFFFFF1u && g[D - 1] == (WT)0xFFFFFFF3u)   // diagnostic (bit 16): no stores (the test keeps the values live)
                        store_group<WT, D, OT>(o, g);
                } else {
#pragma unroll
                    for (int t = 0; t < D - 1; ++t)
                        if ((u32)t < cnt) __stcs(o + t, (OT)g[t]);
                }
                o += cnt;
            }
        }
        v = vn;
        left = left_next;
#pragma unroll
        for (int t = 0; t < RW; ++t) e[t] = e2[t];
    }
    if constexpr (D == 6) { if (pc) { __stcs(o, (OT)p0); __stcs(o + 1, (OT)p1); o += 2; } }
    if constexpr (sizeof(OT) == 4 && D == 4) { if (pc) { __stcs(reinterpret_cast<uint4*>(o), make_uint4((u32)p0, (u32)p1, (u32)p2, (u32)p3)); o += 4; } }
    if (left == 1) __stcs(o, (OT)v);  // the value carried out of the last full group
}

// "Window pass", the default form. PERSISTENT: a warp takes 32 consecutive items, walks them in lockstep until its
// slowest lane is done (window_items_direct), then strides to its next 32.
// items[i] = (first slot << 8 | cnt, seed): v = seed is the occurrence on the item's first slot, cnt the number of
// further occurrences of the same chain that the item covers (< SEG). Each lookup yields Phi^1..Phi^D(v); the lane
// emits the aligned group [v, Phi(v), .., Phi^(D-1)(v)] as one 32-byte sector store and continues from Phi^D(v).
template <typename WT, int D, bool KEEP, int MINB, typename OT>
__global__ void __launch_bounds__(256, MINB)
phi_window_batch_kernel(const FlatDev ix, const u64* __restrict__ items, const u64* __restrict__ ctr, OT* __restrict__ out,
                        u64 cap, u64 items_cap, u32 seg_shift) {
    u64 total, total_chains;
    if (!expansion_enabled(ctr, cap, items_cap, seg_shift, true, total, total_chains)) return;
    const u32 n_items = (u32)__ldcg(ctr + RIG_CTR_ITEMS);
    const u32 T = gridDim.x * blockDim.x;
    for (u32 ib = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); ib < n_items; ib += T) {  // warp-uniform
        const u32 i = ib + (threadIdx.x & 31u);
        u32 left = 0;
        OT* o = out;
        WT v = 0;
        if (i < n_items) {
            const ulonglong2 it = __ldcs(reinterpret_cast<const ulonglong2*>(items) + i);
            left = (u32)(it.x & 255u) + 1;
            o = out + (it.x >> 8);
            v = (WT)it.y;
        }
        window_items_direct<WT, D, KEEP, OT>(ix, left, o, v);
    }
}

// FUSED expansion: seed pass and window pass in ONE persistent kernel. Every warp first PRODUCES — its share of the
// chains: toeholds, heads, one item per seed-table hop (produce_chains) — and then CONSUMES items, 32 at a time by
// ticket, as the list fills: the seed pass is a chain of dependent DRAM-latency hops per chain (the longest chain of
// config C2 takes 31) that leaves the machine idle, the window pass is bound by the SM's request port; fused, the
// first hides under the second. A consumer accepts an item when both of its words carry this call's epoch tag
// (produce_chains), and gives up on a ticket only when every warp has finished producing (ctr[RIG_CTR_DONE], a
// device-scope fence on both sides) and the ticket lies beyond the final item count. No consumer ever waits for a
// warp that is not running: the grid is sized to be fully resident.
#define RIG_CTR_DONE 7    // warps that have finished producing
#define RIG_CTR_TAKEN 8   // items handed out to consumers
template <typename WT, int D, bool KEEP, int MINB, typename OT>
__global__ void __launch_bounds__(256, MINB)
phi_fused_kernel(const FlatDev ix, u64 N, const u64* __restrict__ ch_off, const u64* __restrict__ occ_off,
                 const u64* __restrict__ lo_in, const u64* __restrict__ hi_in, const u64* __restrict__ toe_in,
                 const u64* __restrict__ jl_in, OT* __restrict__ out, u64* __restrict__ ctr, u64 cap,
                 u64* __restrict__ items, u64 items_cap, u32 seg_shift, u64 tag, u32 prod_mod) {
    u64 total, total_chains;
    if (!expansion_enabled(ctr, cap, items_cap, seg_shift, true, total, total_chains)) return;
    const u32 lane = threadIdx.x & 31u;
    // Producers: the CTAs with blockIdx % prod_mod == 0 (spread over the SMs); the others consume from the start.
    // Every warp producing would leave nobody to consume until the longest chains are done.
    const u32 n_prod_ctas = (gridDim.x + prod_mod - 1) / prod_mod;
    if (blockIdx.x % prod_mod == 0) {
        produce_chains<WT, D, KEEP, true, OT>(ix, N, ch_off, occ_off, lo_in, hi_in, toe_in, jl_in, out, ctr, items, seg_shift, total_chains, tag,
                                          blockIdx.x / prod_mod, n_prod_ctas);
        __syncwarp();
        if (lane == 0) { __threadfence(); atomicAdd(ctr + RIG_CTR_DONE, 1ull); }
    }
    const u64 nwarps = (u64)n_prod_ctas * (blockDim.x >> 5);
    bool all_done = false;     // every warp has finished producing (warp-uniform)
    u64 n_final = 0;           // then: the final item count
    for (;;) {
        u64 ib = 0;
        if (lane == 0) ib = atomicAdd(ctr + RIG_CTR_TAKEN, 32ull);
        ib = __shfl_sync(RIG_FULL, ib, 0);
        if (ib >= items_cap) break;   // the list never holds more than items_cap entries (expansion_enabled): nothing can appear here
        const u64 i = ib + lane;
        bool have = false;
        u64 w0 = 0, w1 = 0;
        if (!all_done) n_final = items_cap;   // until the final count is known: the bound of what may still be written
        for (;;) {
            if (!have && i < n_final) {
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(items + 2 * i) : "memory");
                have = ((w0 ^ tag) >> 48) == 0 && ((w1 ^ tag) >> 48) == 0;
            }
            const bool waiting = !have && i < n_final;
            if (!__any_sync(RIG_FULL, waiting)) break;
            if (!all_done) {
                u64 d = 0;
                if (lane == 0) d = ld_relaxed_gpu(ctr + RIG_CTR_DONE);
                d = __shfl_sync(RIG_FULL, d, 0);
                if (d == nwarps) {
                    __threadfence();
                    if (lane == 0) n_final = ld_relaxed_gpu(ctr + RIG_CTR_ITEMS);
                    n_final = __shfl_sync(RIG_FULL, n_final, 0);
                    all_done = true;
                    continue;   // every item below n_final is visible now: one more look
                }
                __nanosleep(200);
            }
        }
        if (all_done && ib >= n_final) break;
        u32 left = 0;
        OT* o = out;
        WT v = 0;
        if (have) {
            left = (u32)(w0 & 255u) + 1;
            o = out + ((w0 & RIG_ITEM_MASK) >> 8);
            v = (WT)(w1 & RIG_ITEM_MASK);
        }
        window_items_direct<WT, D, KEEP, OT>(ix, left, o, v);
    }
}

}  // namespace rigk
