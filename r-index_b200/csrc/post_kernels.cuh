// post_kernels.cuh — sm_100a device code for ri-locate's post-processing options:
//
//   -o  (reference ri-locate.cpp:146-152)  per pattern, std::sort of its occurrences, ascending
//       -> segsort_*: one CTA per pattern's segment of the occurrence array, bitonic merge network in
//          shared memory (keys narrowed to 32 bits when n < 2^32), in global memory for the few segments
//          that do not fit.
//   -c  (reference ri-locate.cpp:156-190)  per pattern: count its occurrences in the text by brute force
//       (string::find loop), compare with the number located, and compare text[o, o+m) with the pattern for
//       every located o.
//       -> check_positions_kernel (one thread per occurrence: bytes equal, strictly ascending inside the
//          pattern) and a HASH JOIN for the brute-force counts: the patterns go into an open-addressing
//          table keyed by a 64-bit hash of their m bytes (duplicates share one entry), one thread per text
//          position hashes its m-gram and probes the table. N x n string::find scans become one pass over
//          the text: the reference's self-check at full speed.
#pragma once
#include "search_kernels.cuh"

namespace rigk {

// ------------------------------------------------------------------ segmented sort
// Bitonic sorter in its "all comparisons ascending" form: merging blocks of size k starts with a MIRROR step
// (i <-> block_end - 1 - i), followed by half-cleaners at distances k/4, k/8, .., 1. Because every
// compare-exchange puts the smaller key at the lower index, positions >= len behave as +infinity without
// being stored: a compare-exchange whose upper index is >= len is skipped. No padding to a power of two.
template <typename KT>
__device__ __forceinline__ void bitonic_sort_inplace(KT* s, u32 len, u32 tid, u32 nthreads) {
    u32 P = 1;
    while (P < len) P <<= 1;
    for (u32 k = 2; k <= P; k <<= 1) {
        const u32 hk = k >> 1;
        for (u32 i = tid; i < (P >> 1); i += nthreads) {  // mirror step
            const u32 blk = i / hk, off = i - blk * hk;
            const u32 a = blk * k + off, b = blk * k + k - 1 - off;
            if (b < len) { const KT x = s[a], y = s[b]; if (x > y) { s[a] = y; s[b] = x; } }
        }
        __syncthreads();
        for (u32 j = hk >> 1; j >= 1; j >>= 1) {          // half-cleaners
            for (u32 i = tid; i < (P >> 1); i += nthreads) {
                const u32 a = ((i / j) * (j << 1)) + (i % j), b = a + j;
                if (b < len) { const KT x = s[a], y = s[b]; if (x > y) { s[a] = y; s[b] = x; } }
            }
            __syncthreads();
        }
    }
}

// One CTA per pattern; segments of 2..cap keys are sorted in shared memory (dynamic, cap * sizeof(KT) bytes).
// Longer segments are appended to big[] (count in big_count) for the next tier. `seg_list` != nullptr: the CTA
// takes its pattern index from seg_list[blockIdx.x] (tiers 2 and 3), bounded by *seg_count.
template <typename KT>
__global__ void __launch_bounds__(1024)
segsort_smem_kernel(const u64* __restrict__ occ_off, u64* __restrict__ occ, u64 N, u32 cap,
                    const u32* __restrict__ seg_list, const u64* __restrict__ seg_count,
                    u32* __restrict__ big, u64* __restrict__ big_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KT* s = reinterpret_cast<KT*>(smem_raw);
    u64 p = blockIdx.x;
    if (seg_list) {
        if (p >= __ldcg(seg_count)) return;
        p = seg_list[p];
    }
    if (p >= N) return;
    const u64 a = occ_off[p], b = occ_off[p + 1];
    const u64 len64 = b - a;
    if (len64 < 2) return;
    if (len64 > cap) {
        if (threadIdx.x == 0 && big) big[atomicAdd(big_count, 1ull)] = (u32)p;
        return;
    }
    const u32 len = (u32)len64;
    u64* g = occ + a;
    for (u32 i = threadIdx.x; i < len; i += blockDim.x) s[i] = (KT)g[i];
    __syncthreads();
    bitonic_sort_inplace<KT>(s, len, threadIdx.x, blockDim.x);
    for (u32 i = threadIdx.x; i < len; i += blockDim.x) g[i] = (u64)s[i];
}

// Last tier: segments of any length, sorted in place in global memory (L2-resident for the sizes that occur)
// by one CTA each. Lengths >= 2^32 are not supported (a pattern cannot have more than n < 2^63 occurrences,
// but one CTA would not finish; the host refuses them).
__global__ void __launch_bounds__(1024)
segsort_global_kernel(const u64* __restrict__ occ_off, u64* __restrict__ occ, const u32* __restrict__ seg_list,
                      const u64* __restrict__ seg_count) {
    if (blockIdx.x >= __ldcg(seg_count)) return;
    const u64 p = seg_list[blockIdx.x];
    const u64 a = occ_off[p], b = occ_off[p + 1];
    bitonic_sort_inplace<u64>(occ + a, (u32)(b - a), threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------------ -c check
struct CheckReport {  // mirrors rig_check_report (include/rindex_gpu.h)
    u64 patterns_checked, wrong_count_patterns, wrong_occurrences, unsorted_or_duplicate, first_bad_pattern,
        first_bad_position;
};

__device__ __forceinline__ u64 hash_bytes(const uint8_t* p, u64 m) {  // FNV-1a, finalised (table index = low bits)
    u64 h = 1469598103934665603ull;
    for (u64 i = 0; i < m; ++i) { h ^= (u64)__ldg(p + i); h *= 1099511628211ull; }
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return h;
}

__device__ __forceinline__ bool bytes_equal(const uint8_t* a, const uint8_t* b, u64 m) {
    for (u64 i = 0; i < m; ++i) if (__ldg(a + i) != __ldg(b + i)) return false;
    return true;
}

// table[slot] = pattern index + 1 (0 = empty); rep[p] = the first-inserted pattern with the same bytes
__global__ void __launch_bounds__(256)
check_build_kernel(const uint8_t* __restrict__ patt, u64 N, u64 m, u32* __restrict__ table, u64 mask,
                   u32* __restrict__ rep) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const uint8_t* P = patt + p * m;
    u64 s = hash_bytes(P, m) & mask;
    for (;;) {
        u32 cur = atomicCAS(table + s, 0u, (u32)p + 1u);
        if (cur == 0u) { rep[p] = (u32)p; return; }
        if (bytes_equal(patt + (u64)(cur - 1u) * m, P, m)) { rep[p] = cur - 1u; return; }
        s = (s + 1) & mask;
    }
}

// one thread per text position i (0 <= i <= len - m): found[q] += 1 for the table entry q whose bytes equal text[i, i+m)
__global__ void __launch_bounds__(256)
check_scan_kernel(const uint8_t* __restrict__ text, u64 len, const uint8_t* __restrict__ patt, u64 m,
                  const u32* __restrict__ table, u64 mask, u64* __restrict__ found) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (m == 0 || len < m || i > len - m) return;
    const uint8_t* T = text + i;
    u64 s = hash_bytes(T, m) & mask;
    for (;;) {
        const u32 cur = __ldg(table + s);
        if (cur == 0u) return;
        if (bytes_equal(patt + (u64)(cur - 1u) * m, T, m)) { atomicAdd(found + (cur - 1u), 1ull); return; }
        s = (s + 1) & mask;
    }
}

// per pattern: brute-force count vs located count (ri-locate.cpp:170-176)
__global__ void __launch_bounds__(256)
check_counts_kernel(const u64* __restrict__ lo, const u64* __restrict__ hi, const u32* __restrict__ rep,
                    const u64* __restrict__ found, u64 N, u64 m, u64 text_len, CheckReport* __restrict__ rp) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const u64 want = hi[p] >= lo[p] ? hi[p] - lo[p] + 1 : 0;
    // m = 0: the reference's find loop is not meaningful; the full range has n = text_len + 1 rows
    const u64 got = m == 0 ? text_len + 1 : found[rep[p]];
    if (want != got) {
        atomicAdd(&rp->wrong_count_patterns, 1ull);
        atomicMin(&rp->first_bad_pattern, p);
    }
}

// one thread per occurrence: text[o, o+m) == pattern (ri-locate.cpp:178-188), and strictly ascending inside the
// pattern (sorted + distinct: the located set has no repeats)
__global__ void __launch_bounds__(256)
check_positions_kernel(const uint8_t* __restrict__ text, u64 len, const uint8_t* __restrict__ patt, u64 N, u64 m,
                       const u64* __restrict__ occ_off, const u64* __restrict__ occ, u64 total, int sorted,
                       CheckReport* __restrict__ rp) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    u64 a = 0, b = N;  // largest p with occ_off[p] <= i
    while (b - a > 1) {
        const u64 mid = (a + b) >> 1;
        if (__ldg(occ_off + mid) <= i) a = mid; else b = mid;
    }
    const u64 p = a, o = occ[i];
    bool ok = (m == 0) ? (o <= len) : (o + m <= len && bytes_equal(text + o, patt + p * m, m));
    if (!ok) {
        atomicAdd(&rp->wrong_occurrences, 1ull);
        atomicMin(&rp->first_bad_pattern, p);
        atomicMin(&rp->first_bad_position, o);
    }
    if (sorted && i > __ldg(occ_off + p) && occ[i - 1] >= o) atomicAdd(&rp->unsorted_or_duplicate, 1ull);
}

}  // namespace rigk
