// phi_kernels.cuh — sm_100a Phi expansion: the locate_all loop of the reference
// (internal/r_index.hpp:340-351: OCC = SA[hi], Phi(SA[hi]), Phi^2(SA[hi]), ...) for a whole batch.
//
// TWO PASSES when the index carries a seed table (Phi^SEG, flat_layout.hpp: JumpTable), one otherwise:
//
//   phi_expand_kernel<.., SEEDED=true>   "seed pass", one lane per chain: writes the toehold, walks the chain's
//        head up to the next 128-byte line of the OUTPUT array (<= 16 occurrences), then hops along the chain
//        SEG occurrences at a time (one seed-table lookup per hop) and appends one 16-byte ITEM per hop to a
//        list: (first output slot, number of occurrences that follow, seed = the occurrence on that slot).
//   phi_window_kernel                    "window pass", one lane per item: from the seed, every lookup yields D
//        occurrences that leave as one aligned 32-byte sector.
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
#include "search_kernels.cuh"

namespace rigk {

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

template <typename WT, int D>
__device__ __forceinline__ void store_group(u64* p, const WT (&e)[D]) {  // p is D*8-byte aligned
    if constexpr (D == 1) __stcs(p, (u64)e[0]);
    else if constexpr (D == 2) __stcs(reinterpret_cast<ulonglong2*>(p), make_ulonglong2((u64)e[0], (u64)e[1]));
    else {
#pragma unroll
        for (int k = 0; k < D / 4; ++k)
            stg256_stream(p + 4 * k, (u64)e[4 * k], (u64)e[4 * k + 1], (u64)e[4 * k + 2], (u64)e[4 * k + 3]);
    }
}

// Pull the Phi tables into L2 with one streaming pass (evict_last) before the walk: a cold table costs
// every chain a DRAM round trip per first touch, and with 32 lanes in lockstep almost every iteration
// of every warp would contain one. ~40 MB at HBM speed is a few microseconds.
__global__ void __launch_bounds__(256) l2_warm_kernel(const char* base, u64 bytes) {
    const u64 line = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    if (line < bytes) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(base + line));
}

// ---- walking a chain with the Phi^1..Phi^D table ---------------------------------------------------
// From v = SA[x] (already written at o[-1]) produce `remaining` further occurrences SA[x-1], SA[x-2], ...
// at o[0], o[1], ...; returns the last value produced (v itself when remaining == 0).
//
// Per step, from the current value v = SA[x]: e[t] = Phi^(t+1)(v) = (v + delta_t) mod n, where the
// deltas are those of the piece holding v. For t = 0 this is r_index::Phi (r_index.hpp:195-221):
// strict circular predecessor over the sorted run-first samples (sparse_sd_vector.hpp:107-112,153-157)
// and (prev_sample + delta) % n (:219), folded into one delta per piece.
template <typename WT, int D, bool KEEP>
__device__ __forceinline__ WT walk_chain(const FlatDev& ix, WT v, u64* o, WT remaining) {
    constexpr int RW = (D == 1) ? 4 : ((D <= 4) ? 8 : 16);
    const u32 ESZ = ix.phi.esz;  // entry size in bytes (RW words, or 32 when packed)
    const bool PK = ix.phi.packed != 0;
    constexpr bool W32 = sizeof(WT) == 4;
    const WT n = (WT)ix.n;
    const char* rec = reinterpret_cast<const char*>(ix.phi.rec);
    const char* pent = reinterpret_cast<const char*>(ix.phi.pent);
    const u32 shift = ix.phi.shift;
    // slots up to the next D*8-byte boundary go out singly; afterwards every emit is one aligned group
    u32 take = D - (u32)((reinterpret_cast<unsigned long long>(o) >> 3) % D);
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
            slo = (u32)e[D + 1]; shi = slo + (u32)e[D + 2] - 1;
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
            if (cnt == (u32)D) {
                if (!(ix.pad & 1)) store_group<WT, D>(o, x);  // pad bit0: diagnostic "no stores" run (RIG_VARIANT bit4)
            } else {
#pragma unroll
                for (int t = 0; t < D - 1; ++t)
                    if ((u32)t < cnt) __stcs(o + t, (u64)x[t]);
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

#define RIG_LINE 16  // output slots per 128-byte line

// Work item w -> (pattern p, run j): BWT positions [max(lo,start[j]), min(hi,start[j+1]-1)], walked
// from the top down. Output slot of SA[x] is occ_off[p] + (hi - x): locate_all order (r_index.hpp:340-351).
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
template <typename WT, int D, bool KEEP, bool SEEDED>
__global__ void __launch_bounds__(256)
phi_expand_kernel(const FlatDev ix, u64 N, const u64* __restrict__ ch_off, const u64* __restrict__ occ_off,
                  const u64* __restrict__ lo_in, const u64* __restrict__ hi_in, const u64* __restrict__ toe_in,
                  const u64* __restrict__ jl_in, u64* __restrict__ out, u64 total_chains,
                  u64* __restrict__ items, u64* __restrict__ item_count, u32 seg_shift) {
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = w < total_chains;
    if (!SEEDED && !active) return;
    u64 g0 = 0, glast = 0, v0 = 0;
    if (active) {
        u64 a = 0, b = N;  // largest p with ch_off[p] <= w
        while (b - a > 1) {
            const u64 mid = (a + b) >> 1;
            if (__ldg(ch_off + mid) <= w) a = mid; else b = mid;
        }
        const u64 p = a;
        const u64 L = __ldg(lo_in + p), H = __ldg(hi_in + p);
        const u64 j = __ldg(jl_in + p) + (w - __ldg(ch_off + p));
        const u64 sj = ld_pos<WT>(ix.start, j), ej = (u64)ld_pos<WT>(ix.start, j + 1) - 1;
        const u64 top = min(H, ej), bot = max(L, sj);
        if (top == H) v0 = __ldg(toe_in + p);  // toehold carried by the backward search (r_index.hpp:482-545)
        else { v0 = (u64)ld_pos<WT>(ix.samples_last, j) + 1; if (v0 >= ix.n) v0 -= ix.n; }  // run end: SA = sample + 1
        g0 = __ldg(occ_off + p) + (H - top);  // slot of the chain's first occurrence
        glast = g0 + (top - bot);             // slot of its last (a chain never exceeds n)
        __stcs(out + g0, v0);
    }
    if (!SEEDED) {
        walk_chain<WT, D, KEEP>(ix, (WT)v0, out + g0 + 1, (WT)(glast - g0));
    } else {
        const u64 SEG = 1ull << seg_shift;
        const u64 a1 = (g0 + RIG_LINE - 1) & ~(u64)(RIG_LINE - 1);  // first line-aligned slot at or after g0
        const u64 K = (active && a1 <= glast) ? ((glast - a1) >> seg_shift) + 1 : 0;  // items of this chain
        // reserve K entries of items[]: inclusive warp scan, one atomic by the last lane
        const int lane = threadIdx.x & 31;
        u64 incl = K;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t = __shfl_up_sync(RIG_FULL, incl, d);
            if (lane >= d) incl += t;
        }
        u64 wbase = 0;
        if (lane == 31 && incl) wbase = atomicAdd(item_count, incl);
        wbase = __shfl_sync(RIG_FULL, wbase, 31);
        if (!active) return;
        const u64 pre_last = min(a1, glast);
        WT v = walk_chain<WT, D, KEEP>(ix, (WT)v0, out + g0 + 1, (WT)(pre_last - g0));
        if (K) {
            ulonglong2* it = reinterpret_cast<ulonglong2*>(items) + wbase + (incl - K);
            u64 s = a1;
            for (;;) {
                __stcs(it++, make_ulonglong2((s << 8) | min(SEG - 1, glast - s), (u64)v));
                s += SEG;
                if (s > glast) break;
                v = seed_hop<WT>(ix, v);
            }
        }
    }
}

// pair access to a staging row (8-byte pairs of u32, 16-byte pairs of u64)
__device__ __forceinline__ void st_pair(u32* row, u32 pair, u32 a, u32 b) { reinterpret_cast<uint2*>(row)[pair] = make_uint2(a, b); }
__device__ __forceinline__ void st_pair(u64* row, u32 pair, u64 a, u64 b) { reinterpret_cast<ulonglong2*>(row)[pair] = make_ulonglong2(a, b); }
__device__ __forceinline__ void ld_pair(const u32* row, u32 pair, u64& a, u64& b) { const uint2 x = reinterpret_cast<const uint2*>(row)[pair]; a = x.x; b = x.y; }
__device__ __forceinline__ void ld_pair(const u64* row, u32 pair, u64& a, u64& b) { const ulonglong2 x = reinterpret_cast<const ulonglong2*>(row)[pair]; a = x.x; b = x.y; }

__device__ __forceinline__ void stg128_stream(void* p, u64 a, u64 b) {
    // (the .L2::evict_first qualifier is only accepted on 256-bit stores; .cs is the 128-bit streaming form)
    asm volatile("st.global.cs.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}

// One lane per item: items[i] = (first slot << 8 | cnt, seed): v = seed is the occurrence on the item's first
// slot, cnt the number of further occurrences of the same chain that the item covers (< SEG). Each lookup
// yields Phi^1..Phi^D(v); the lane emits the aligned group [v, Phi(v), .., Phi^(D-1)(v)] and continues from
// Phi^D(v). Same per-lane state machine and software pipelining as walk_chain.
//
// STORES. A lane's groups are 32-byte sectors of its own 128-byte lines; stored directly (STAGE = false, the
// default) they cost one L2 request per sector, and the kernel runs at the L2 tag-lookup rate
// (lts__t_tag_requests 80%) and the SM's request port (l1tex2xbar 72%: one load request or one 32-byte store
// payload per cycle), not at a byte rate. STAGE = true is the measured alternative (RIG_VARIANT bit 6): every
// line that will be complete is staged in shared memory (two rows per lane, pair-swizzled; a lane emits at
// most one group per trip, so between two flushes it completes at most one row and starts the other) and
// every FLUSH_TRIPS trips the warp writes out the completed rows, 8 lanes x 16 bytes per line, 4 whole
// lines per store instruction. That halves the tag requests (80% -> 37%) but the store payload through the
// request port is unchanged and the extra instructions (+70%) make it issue-bound: 0.33 ms vs 0.27 ms on C2,
// 1.93 vs 1.79 ms on C3s. Kept for the record, off by default.
template <typename WT, int D, bool KEEP, bool STAGE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
phi_window_kernel(const FlatDev ix, const u64* __restrict__ items, const u64* __restrict__ item_count,
                  u64* __restrict__ out) {
    constexpr int RW = (D == 1) ? 4 : ((D <= 4) ? 8 : 16);
    const u32 ESZ = ix.phi.esz;
    const bool PK = ix.phi.packed != 0;
    constexpr bool W32 = sizeof(WT) == 4;
    constexpr int GPL = RIG_LINE / D;  // groups per line
    // a lane emits at most one group per trip: with a flush every <= GPL trips it cannot complete a second row
    // while one is pending (4 trips = one row per lane in lockstep for D = 4)
    constexpr u32 FLUSH_TRIPS = GPL < 4 ? GPL : 4;
    static_assert(FLUSH_TRIPS <= GPL && (FLUSH_TRIPS & (FLUSH_TRIPS - 1)) == 0, "flush period");
    __shared__ __align__(16) WT stage[STAGE ? WARPS : 1][STAGE ? 2 : 1][STAGE ? 32 : 1][RIG_LINE];
    __shared__ uint8_t sidx[STAGE ? WARPS : 1][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 sw = lane & 7;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 n_items = __ldcg(item_count);
    u32 left = 0;  // slots of this item still to be written, the seed's included
    u64* o = out;
    WT v = 0;
    if (i < n_items) {
        const ulonglong2 it = __ldcs(reinterpret_cast<const ulonglong2*>(items) + i);
        left = (u32)(it.x & 255u) + 1;
        o = out + (it.x >> 8);
        v = (WT)it.y;
    }
    const WT n = (WT)ix.n;
    const char* rec = reinterpret_cast<const char*>(ix.phi.rec);
    const char* pent = reinterpret_cast<const char*>(ix.phi.pent);
    const u32 shift = ix.phi.shift;
    bool searching = false, staging = false, pending = false;  // pending: row cur^1 is complete, not yet written out
    u32 slo = 0, shi = 0, probe = 0, fill = 0, cur = 0, trip = 0;  // fill: groups staged in row cur
    u64* pend_line = out;
    WT e[RW];
    if (left > 1) load_entry<WT, RW, KEEP>(rec + (u64)(v >> shift) * ESZ, e, PK);
    for (;;) {
        const bool more = __any_sync(RIG_FULL, left > 1);
        // ---- write out completed rows: every FLUSH_TRIPS trips and once at the end (warp-uniform) ----
        if (STAGE && (!more || (trip & (FLUSH_TRIPS - 1)) == FLUSH_TRIPS - 1)) {
            const u32 ready = __ballot_sync(RIG_FULL, pending);
            if (ready) {
                if (pending) sidx[wid][__popc(ready & ((1u << lane) - 1u))] = (uint8_t)(lane | ((cur ^ 1u) << 5));
                __syncwarp();
                const u32 nready = __popc(ready);
                for (u32 t = 0; t < nready; t += 4) {
                    const u32 k = t + (lane >> 3);
                    const bool act = k < nready;
                    const u32 sx = act ? sidx[wid][k] : 0u;
                    const u32 src = sx & 31u;
                    const u64 dst = __shfl_sync(RIG_FULL, (unsigned long long)pend_line, src);
                    if (act && !(ix.pad & 1)) {
                        u64 x0, x1;
                        ld_pair(&stage[wid][sx >> 5][src][0], sw ^ (src & 7), x0, x1);  // slots 2*sw, 2*sw+1 of the line
                        stg128_stream(reinterpret_cast<u64*>(dst) + 2 * sw, x0, x1);
                    }
                }
                __syncwarp();
                pending = false;
            }
        }
        if (!more) break;
        ++trip;
        bool emit = false;
        WT g[D];          // the group to store: [v, Phi(v), ..]
        u32 cnt = 0;
        WT vn = v;
        if (left > 1) {
            if (!searching) {
                emit = v < e[D];
                slo = (u32)e[D + 1]; shi = slo + (u32)e[D + 2] - 1;
                searching = !emit;
            } else if (slo == shi) {
                emit = true;
            } else if (e[D] <= v) {
                slo = probe; emit = (slo == shi);
            } else {
                shi = probe - 1; emit = false;
            }
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
        if (left_next > 1)
            load_entry<WT, RW, KEEP>(searching ? (pent + (u64)probe * ESZ) : (rec + (u64)(vn >> shift) * ESZ), e2, PK);
        if (emit) {
            // stage the line iff all of its GPL groups will be emitted as full groups (with D = 1 the loop ends
            // at left == 1, one slot early, so one more slot is needed)
            if (STAGE && fill == 0) staging = left >= (u32)(RIG_LINE + (D == 1 ? 1 : 0));
            if (cnt == (u32)D) {
                if (STAGE && staging) {
                    WT* row = &stage[wid][cur][lane][0];
                    if constexpr (D == 1) {
                        row[(((fill >> 1) ^ sw) << 1) | (fill & 1)] = g[0];
                    } else {
#pragma unroll
                        for (int t = 0; t < D; t += 2)
                            st_pair(row, ((fill * D + t) >> 1) ^ sw, g[t], g[t + 1]);
                    }
                    if (++fill == (u32)GPL) {  // row complete: hand it to the next flush, continue in the other row
                        fill = 0; cur ^= 1u; pending = true;
                        pend_line = o + D - RIG_LINE;
                    }
                } else if (!(ix.pad & 1)) {
                    store_group<WT, D>(o, g);
                }
            } else {
#pragma unroll
                for (int t = 0; t < D - 1; ++t)
                    if ((u32)t < cnt) __stcs(o + t, (u64)g[t]);
            }
            o += cnt;
        }
        v = vn;
        left = left_next;
#pragma unroll
        for (int t = 0; t < RW; ++t) e[t] = e2[t];
    }
    if (left == 1) __stcs(o, (u64)v);  // the value carried out of the last full group
}

}  // namespace rigk
