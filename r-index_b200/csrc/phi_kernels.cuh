// phi_kernels.cuh — sm_100a Phi expansion: the locate_all loop of the reference
// (internal/r_index.hpp:340-351: OCC = SA[hi], Phi(SA[hi]), Phi^2(SA[hi]), ...) for a whole batch.
//
// One LANE per chain. A chain is the part of a pattern's SA range that lies inside one BWT run
// (ranges are cut at run boundaries: every run end is a free toehold, SA = samples_last[run]+1, the
// identity behind r_index.hpp:489,533), so a batch exposes (#patterns x #runs overlapped) independent
// chains instead of #patterns.
//
// One record LOOKUP yields D occurrences: the PhiTable (flat_layout.hpp) stores, per piece of the
// refined translation, the deltas of Phi^1..Phi^D, so from SA[x] one 16/32/64-byte bucket record gives
// SA[x-1..x-D]. D consecutive outputs of a chain are D consecutive u64 slots: they leave as ONE
// aligned vector store (256-bit for D = 4). Per 4 occurrences the kernel issues 1 random 32-byte
// sector load + 1 full-sector store — the measured limiter of the previous versions was L1/TEX
// sector throughput (one divergent sector per occurrence), see DESIGN.md §5.
#pragma once
#include "search_kernels.cuh"

namespace rigk {

__device__ __forceinline__ void ldg256(const void* p, u64& a, u64& b, u64& c, u64& d) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void stg256_stream(void* p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.u64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// Load `NW` consecutive table words starting at word index `idx` into w[] (zero-extended to u64).
// Addresses are NW*wordsize aligned by construction.
template <bool W32, int NW>
__device__ __forceinline__ void load_words(const void* base, u64 idx, u64 (&w)[NW]) {
    if (W32) {
        const u32* p = reinterpret_cast<const u32*>(base) + idx;
        if constexpr (NW == 1) { w[0] = __ldg(p); }
        else if constexpr (NW == 2) { const uint2 x = __ldg(reinterpret_cast<const uint2*>(p)); w[0] = x.x; w[1] = x.y; }
        else if constexpr (NW == 4) { const uint4 x = __ldg(reinterpret_cast<const uint4*>(p)); w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; }
        else {
#pragma unroll
            for (int k = 0; k < NW / 8; ++k) {
                u64 a, b, c, d;
                ldg256(p + 8 * k, a, b, c, d);
                w[8 * k + 0] = (u32)a; w[8 * k + 1] = a >> 32; w[8 * k + 2] = (u32)b; w[8 * k + 3] = b >> 32;
                w[8 * k + 4] = (u32)c; w[8 * k + 5] = c >> 32; w[8 * k + 6] = (u32)d; w[8 * k + 7] = d >> 32;
            }
        }
    } else {
        const u64* p = reinterpret_cast<const u64*>(base) + idx;
        if constexpr (NW == 1) { w[0] = __ldg(p); }
        else if constexpr (NW == 2) { const ulonglong2 x = __ldg(reinterpret_cast<const ulonglong2*>(p)); w[0] = x.x; w[1] = x.y; }
        else {
#pragma unroll
            for (int k = 0; k < NW / 4; ++k) ldg256(p + 4 * k, w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
        }
    }
}

// e[j] = Phi^(j+1)(i), j = 0..D-1. For j = 0 this is r_index::Phi (r_index.hpp:195-221): strict
// circular predecessor over the sorted run-first samples (sparse_sd_vector.hpp:107-112,153-157) and
// (prev_sample + delta) % n (:219), folded into one delta per piece.
template <bool W32, int D>
__device__ __forceinline__ void phi_lookup(const PhiTabDev& T, u64 i, u64 n, u64 (&e)[D]) {
    constexpr int RW = (D == 1) ? 4 : ((D <= 4) ? 8 : 16);
    const u64 q = i >> T.shift;
    u64 w[RW];
    load_words<W32, RW>(T.rec, q * RW, w);
    u64 d[D];
#pragma unroll
    for (int j = 0; j < D; ++j) d[j] = w[j];
    if (i >= w[D]) {  // a piece starts inside the bucket at or below i (w[D] = ~0 when none does)
        u64 piece = w[D + 1];
        if (i >= w[D + 2]) {  // rare: two or more do -> last piece in (piece, dir[q+1]] with start <= i
            u64 lo = piece + 1, hi = __ldg(T.dir + q + 1);
            while (lo < hi) {
                const u64 mid = (lo + hi + 1) >> 1;
                if (__ldg(T.start + mid) <= i) lo = mid; else hi = mid - 1;
            }
            piece = lo;
        }
        load_words<W32, D>(T.delta, piece * D, d);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
        u64 v = i + d[j];
        if (v >= n) v -= n;
        e[j] = v;
    }
}

template <int D>
__device__ __forceinline__ void store_group(u64* p, const u64 (&e)[D]) {  // p is D*8-byte aligned
    if constexpr (D == 1) __stcs(p, e[0]);
    else if constexpr (D == 2) __stcs(reinterpret_cast<ulonglong2*>(p), make_ulonglong2(e[0], e[1]));
    else {
#pragma unroll
        for (int k = 0; k < D / 4; ++k) stg256_stream(p + 4 * k, e[4 * k], e[4 * k + 1], e[4 * k + 2], e[4 * k + 3]);
    }
}

// Work item w -> (pattern p, run j): BWT positions [max(lo,start[j]), min(hi,start[j+1]-1)], walked
// from the top down. Output slot of SA[x] is occ_off[p] + (hi - x): locate_all order (r_index.hpp:340-351).
template <bool W32, int D>
__global__ void __launch_bounds__(256)
phi_expand_kernel(const FlatDev ix, u64 N, const u64* __restrict__ ch_off, const u64* __restrict__ occ_off,
                  const u64* __restrict__ lo_in, const u64* __restrict__ hi_in, const u64* __restrict__ toe_in,
                  const u64* __restrict__ jl_in, u64* __restrict__ out, u64 total_chains) {
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total_chains) return;  // no warp collectives below
    u64 a = 0, b = N;  // largest p with ch_off[p] <= w
    while (b - a > 1) {
        const u64 mid = (a + b) >> 1;
        if (__ldg(ch_off + mid) <= w) a = mid; else b = mid;
    }
    const u64 p = a;
    const u64 L = __ldg(lo_in + p), H = __ldg(hi_in + p);
    const u64 j = __ldg(jl_in + p) + (w - __ldg(ch_off + p));
    const u64 sj = __ldg(ix.start + j), ej = __ldg(ix.start + j + 1) - 1;
    const u64 top = min(H, ej), bot = max(L, sj);
    u64 v;
    if (top == H) v = __ldg(toe_in + p);  // toehold carried by the backward search (r_index.hpp:482-545)
    else { v = __ldg(ix.samples_last + j) + 1; if (v >= ix.n) v -= ix.n; }  // run end: SA = sample + 1
    const u64 g0 = __ldg(occ_off + p) + (H - top);
    u64* o = out + g0;           // next slot to write
    __stcs(o, v);
    ++o;
    u64 remaining = top - bot;   // occurrences still to produce
    u64 e[D];
    // head: single slots until the next slot index is a multiple of D (vector stores must be aligned)
    const u32 mis = (u32)((g0 + 1) % D);
    if (D > 1 && mis != 0 && remaining > 0) {
        const u32 cnt = (u32)min((u64)(D - mis), remaining);
        phi_lookup<W32, D>(ix.phi, v, ix.n, e);
#pragma unroll
        for (int t = 0; t < D - 1; ++t)
            if ((u32)t < cnt) { __stcs(o + t, e[t]); v = e[t]; }
        o += cnt;
        remaining -= cnt;
    }
    // body: D occurrences per lookup, one aligned vector store
    while (remaining >= D) {
        phi_lookup<W32, D>(ix.phi, v, ix.n, e);
        store_group<D>(o, e);
        v = e[D - 1];
        o += D;
        remaining -= D;
    }
    // tail
    if (D > 1 && remaining > 0) {
        phi_lookup<W32, D>(ix.phi, v, ix.n, e);
#pragma unroll
        for (int t = 0; t < D - 1; ++t)
            if ((u64)t < remaining) __stcs(o + t, e[t]);
    }
}

}  // namespace rigk
