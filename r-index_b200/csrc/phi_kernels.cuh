// phi_kernels.cuh — sm_100a Phi expansion: the locate_all loop of the reference
// (internal/r_index.hpp:340-351: OCC = SA[hi], Phi(SA[hi]), Phi^2(SA[hi]), ...) for a whole batch.
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
// searches a crowded bucket). Measured reason: with an if/else "slow path", one lane
// of 32 taking it stalls the whole warp for 2-6 extra dependent L2 round trips on almost every
// iteration; with the state machine a lane's slow path costs that lane extra iterations only.
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

// Load the RW-word entry at `p` (RW*wordsize aligned) into w[] (zero-extended to u64).
template <bool W32, int RW>
__device__ __forceinline__ void load_entry(const void* p, u64 (&w)[RW]) {
    if (W32) {
        if constexpr (RW == 4) {
            const uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
            w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
        } else {
#pragma unroll
            for (int k = 0; k < RW / 8; ++k) {
                u64 a, b, c, d;
                ldg256(reinterpret_cast<const u32*>(p) + 8 * k, a, b, c, d);
                w[8 * k + 0] = (u32)a; w[8 * k + 1] = a >> 32; w[8 * k + 2] = (u32)b; w[8 * k + 3] = b >> 32;
                w[8 * k + 4] = (u32)c; w[8 * k + 5] = c >> 32; w[8 * k + 6] = (u32)d; w[8 * k + 7] = d >> 32;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < RW / 4; ++k)
            ldg256(reinterpret_cast<const u64*>(p) + 4 * k, w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
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
//
// Per step, from the current value v = SA[x]: e[t] = Phi^(t+1)(v) = (v + delta_t) mod n, where the
// deltas are those of the piece holding v. For t = 0 this is r_index::Phi (r_index.hpp:195-221):
// strict circular predecessor over the sorted run-first samples (sparse_sd_vector.hpp:107-112,153-157)
// and (prev_sample + delta) % n (:219), folded into one delta per piece.
template <bool W32, int D>
__global__ void __launch_bounds__(256)
phi_expand_kernel(const FlatDev ix, u64 N, const u64* __restrict__ ch_off, const u64* __restrict__ occ_off,
                  const u64* __restrict__ lo_in, const u64* __restrict__ hi_in, const u64* __restrict__ toe_in,
                  const u64* __restrict__ jl_in, u64* __restrict__ out, u64 total_chains) {
    constexpr int RW = (D == 1) ? 4 : ((D <= 4) ? 8 : 16);
    constexpr u64 ESZ = (u64)RW * (W32 ? 4 : 8);  // entry size in bytes
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
    u64* o = out + __ldg(occ_off + p) + (H - top);  // next slot to write
    __stcs(o, v);
    ++o;
    u64 remaining = top - bot;  // occurrences still to produce
    const u64 n = ix.n;
    const char* rec = reinterpret_cast<const char*>(ix.phi.rec);
    const char* pent = reinterpret_cast<const char*>(ix.phi.pent);
    const u32 shift = ix.phi.shift;
    bool searching = false;   // false: next load = bucket record of v; true: next load = piece entry `probe`
    u64 slo = 0, shi = 0;     // search interval of piece indices, invariant start[slo] <= v
    while (remaining > 0) {
        const u64 probe = (slo < shi) ? ((slo + shi + 1) >> 1) : slo;
        const char* addr = searching ? (pent + probe * ESZ) : (rec + (v >> shift) * ESZ);
        u64 e[RW];
        load_entry<W32, RW>(addr, e);  // the ONE load of this iteration
        bool emit;
        if (!searching) {
            emit = v < e[D];              // no piece begins inside the bucket at or below v
            if (!emit) { slo = e[D + 1]; shi = slo + e[D + 2] - 1; searching = true; }
        } else if (slo == shi) {
            emit = true;                  // the entry just loaded is the answer
        } else if (e[D] <= v) {
            slo = probe; emit = (slo == shi);
        } else {
            shi = probe - 1; emit = false;
        }
        if (emit) {
            searching = false;
#pragma unroll
            for (int t = 0; t < D; ++t) {
                u64 x = v + e[t];
                if (x >= n) x -= n;
                e[t] = x;
            }
            // slots until the next D*8-byte boundary; full aligned groups leave as one vector store
            const u32 mis = (u32)((reinterpret_cast<unsigned long long>(o) >> 3) % D);
            const u64 cnt = min((u64)(D - mis), remaining);
            if (cnt == D) {
                u64 g[D];
#pragma unroll
                for (int t = 0; t < D; ++t) g[t] = e[t];
                store_group<D>(o, g);
                v = e[D - 1];
            } else {
#pragma unroll
                for (int t = 0; t < D - 1; ++t)
                    if ((u64)t < cnt) { __stcs(o + t, e[t]); v = e[t]; }
            }
            o += cnt;
            remaining -= cnt;
        }
    }
}

}  // namespace rigk
