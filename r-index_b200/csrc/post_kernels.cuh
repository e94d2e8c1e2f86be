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
template <typename KT, typename LT>   // LT: index type (u32 for the shared-memory tiers, u64 in global memory: a segment may exceed 2^31 keys)
__device__ __forceinline__ void bitonic_sort_inplace(KT* s, LT len, LT tid, LT nthreads) {
    LT P = 1;
    while (P < len) P <<= 1;
    for (LT k = 2; k <= P && k != 0; k <<= 1) {
        const LT hk = k >> 1;
        for (LT i = tid; i < (P >> 1); i += nthreads) {  // mirror step
            const LT blk = i / hk, off = i - blk * hk;
            const LT a = blk * k + off, b = blk * k + k - 1 - off;
            if (b < len) { const KT x = s[a], y = s[b]; if (x > y) { s[a] = y; s[b] = x; } }
        }
        __syncthreads();
        for (LT j = hk >> 1; j >= 1; j >>= 1) {          // half-cleaners
            for (LT i = tid; i < (P >> 1); i += nthreads) {
                const LT a = ((i / j) * (j << 1)) + (i % j), b = a + j;
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
    bitonic_sort_inplace<KT, u32>(s, len, threadIdx.x, blockDim.x);
    for (u32 i = threadIdx.x; i < len; i += blockDim.x) g[i] = (u64)s[i];
}

// Last tier: segments of any length, sorted in place in global memory (L2-resident for the sizes that occur)
// by one CTA each, with 64-bit indices (a one-letter pattern on a multi-GB text has more than 2^31 occurrences).
__global__ void __launch_bounds__(1024)
segsort_global_kernel(const u64* __restrict__ occ_off, u64* __restrict__ occ, const u32* __restrict__ seg_list,
                      const u64* __restrict__ seg_count) {
    if (blockIdx.x >= __ldcg(seg_count)) return;
    const u64 p = seg_list[blockIdx.x];
    const u64 a = occ_off[p], b = occ_off[p + 1];
    bitonic_sort_inplace<u64, u64>(occ + a, b - a, (u64)threadIdx.x, (u64)blockDim.x);
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

// ------------------------------------------------------------------ navigation
// Single-position operations of r_index<> as batches, one thread per position (they are off the count /
// locate path; no CLI calls them):
//   NAV_BWT   r_index::operator[](i)  = bwt[i]                      r_index.hpp:162-164, rle_string.hpp:126-131
//   NAV_LF    r_index::LF(i)          = F[c] + bwt.rank(i, c), c = bwt[i]                 r_index.hpp:224-229
//   NAV_FL    r_index::FL(i)          = bwt.select(i - F[c], c), c = F_at(i)              r_index.hpp:232-243
//   NAV_F_AT  r_index::F_at(i)        = upper_bound(F, F + 256, i) - F - 1                r_index.hpp:263-271
enum { NAV_BWT = 0, NAV_LF = 1, NAV_FL = 2, NAV_F_AT = 3 };

// the block holding BWT position x: bdir bucket, then binary search over the blocks' first positions
template <typename PT>
__device__ __forceinline__ u32 nav_block_of(const FlatDev& ix, PT x) {
    const u32 q = (u32)(x >> ix.lf_shift);
    u32 b0 = __ldg(ix.bdir + q), b1 = __ldg(ix.bdir + q + 1);
    while (b0 < b1) {
        const u32 mid = b0 + ((b1 - b0 + 1) >> 1);
        if (ld_pos<PT>(ix.bstart, mid) <= x) b0 = mid; else b1 = mid - 1;
    }
    return b0;
}

template <typename PT>
__device__ __forceinline__ u32 nav_f_at(const FlatDev& ix, u64 i) {  // largest c in [0,255] with F[c] <= i
    u32 lo = 0, hi = 255;
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if (__ldg(ix.F + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename PT>
__global__ void __launch_bounds__(256)
navigate_kernel(const FlatDev ix, int op, const u64* __restrict__ pos, u64 N, u64* __restrict__ out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const u64 i = pos[t];
    if (i >= ix.n) { out[t] = ~0ull; return; }  // outside the BWT: the reference would read out of bounds
    const u32 K = ix.K;
    if (op == NAV_F_AT) { out[t] = nav_f_at<PT>(ix, i); return; }
    if (op == NAV_BWT || op == NAV_LF) {
        const u32 b = nav_block_of<PT>(ix, (PT)i);
        const char* rp = ix.blk + (u64)b * ix.blk_stride;
        const uint8_t* hd = reinterpret_cast<const uint8_t*>(rp + ix.off_head);
        u32 k = 0;  // run of the block holding i: the last one starting at or before i (padding starts are n > i)
        for (u32 g = 1; g < K; ++g) if (rec_word<PT>(ix, rp, 0, g) <= (PT)i) k = g;
        const uint8_t c = __ldg(hd + k);
        if (op == NAV_BWT) { out[t] = c; return; }
        const u32 sidc = __ldg(ix.sid + c);
        u64 rank = rec_word<PT>(ix, rp, ix.off_cum, sidc);  // #c before the block
        for (u32 g = 0; g < k; ++g)
            if (__ldg(hd + g) == c) rank += (u64)(rec_word<PT>(ix, rp, 0, g + 1) - rec_word<PT>(ix, rp, 0, g));
        rank += i - (u64)rec_word<PT>(ix, rp, 0, k);  // #c in bwt[0, i): rle_string::rank(i, c), rle_string.hpp:170-218
        out[t] = __ldg(ix.F + c) + rank;
        return;
    }
    // NAV_FL
    const u32 c = nav_f_at<PT>(ix, i);
    const u64 j = i - __ldg(ix.F + c);  // this c is the j-th (from 0) in column F
    const u32 sidc = __ldg(ix.sid + c);
    // last block whose count of c before it is <= j (rle_string::select, rle_string.hpp:136-165)
    u64 b0 = 0, b1 = ix.nblk - 1;
    while (b0 < b1) {
        const u64 mid = b0 + ((b1 - b0 + 1) >> 1);
        const u64 before = rec_word<PT>(ix, ix.blk + mid * ix.blk_stride, ix.off_cum, sidc);
        if (before <= j) b0 = mid; else b1 = mid - 1;
    }
    const char* rp = ix.blk + b0 * ix.blk_stride;
    const uint8_t* hd = reinterpret_cast<const uint8_t*>(rp + ix.off_head);
    u64 rem = j - (u64)rec_word<PT>(ix, rp, ix.off_cum, sidc);
    u64 res = ~0ull;
    for (u32 g = 0; g < K; ++g) {
        if (__ldg(hd + g) != (uint8_t)c) continue;
        const u64 s0 = rec_word<PT>(ix, rp, 0, g);
        const u64 s1 = (g + 1 < K) ? (u64)rec_word<PT>(ix, rp, 0, g + 1) : (u64)ld_pos<PT>(ix.bstart, b0 + 1);
        if (s0 >= ix.n) break;  // padding
        const u64 len = s1 - s0;
        if (rem < len) { res = s0 + rem; break; }
        rem -= len;
    }
    out[t] = res;
}

// r_index::get_bwt (r_index.hpp:375-377) = rle_string::toString, restricted to a range: bwt[from, from+len) as bytes,
// one thread per output byte group of 16 (binary search for the first run, then a linear walk over run starts).
template <typename PT>
__global__ void __launch_bounds__(256)
bwt_range_kernel(const FlatDev ix, u64 from, u64 len, uint8_t* __restrict__ out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 a = t * 16;
    if (a >= len) return;
    const u64 e = min(len, a + 16);
    u64 x = from + a;
    const u32 b = nav_block_of<PT>(ix, (PT)x);
    u64 run = (u64)b * ix.K;
    while (run + 1 < ix.r && (u64)ld_pos<PT>(ix.start, run + 1) <= x) ++run;
    u64 next = (run + 1 < ix.r) ? (u64)ld_pos<PT>(ix.start, run + 1) : ix.n;
    uint8_t c = __ldg(reinterpret_cast<const uint8_t*>(ix.blk + (run / ix.K) * ix.blk_stride + ix.off_head) + run % ix.K);
    for (u64 o = a; o < e; ++o, ++x) {
        if (x >= next) {
            ++run;
            next = (run + 1 < ix.r) ? (u64)ld_pos<PT>(ix.start, run + 1) : ix.n;
            c = __ldg(reinterpret_cast<const uint8_t*>(ix.blk + (run / ix.K) * ix.blk_stride + ix.off_head) + run % ix.K);
        }
        out[o] = c;
    }
}

// ---- range utilities of the RLBWT as batches ------------------------------------------------
//   RANGE_BREAK_COUNT / RANGE_BREAK_FILL   rle_string::break_range(rn, c)         rle_string.hpp:261-302
//        the maximal sub-ranges of rn = [l, r] that hold only c, given bwt[l] == bwt[r] == c: [l, end of l's run],
//        every c-run strictly between the runs of l and r, [start of r's run, r]; the range itself when l and r
//        lie in one run. A query that breaks the precondition (l > r, r >= n, bwt[l] != c or bwt[r] != c: the
//        reference asserts) yields no range.
//   RANGE_CLOSEST                          rle_string::closest_run_break(rn, c)   rle_string.hpp:455-493
//        bwt[l] == c: the last position of l's run; otherwise the first position >= l holding c (select(rank(l, c), c));
//        ~0 when no c follows l.
// One thread per query walking the run heads between the two runs: O(runs in the range) like the reference's
// O(|result|) ranks and selects on a text with few distinct heads; not tuned (no CLI calls these).
enum { RANGE_BREAK_COUNT = 0, RANGE_BREAK_FILL = 1, RANGE_CLOSEST = 2 };

template <typename PT>
__device__ __forceinline__ u64 nav_run_of(const FlatDev& ix, u64 x) {  // run holding BWT position x < n
    const u32 b = nav_block_of<PT>(ix, (PT)x);
    u64 run = (u64)b * ix.K;
    while (run + 1 < ix.r && (u64)ld_pos<PT>(ix.start, run + 1) <= x) ++run;
    return run;
}
__device__ __forceinline__ uint8_t nav_head(const FlatDev& ix, u64 run) {
    return __ldg(reinterpret_cast<const uint8_t*>(ix.blk + (run / ix.K) * ix.blk_stride + ix.off_head) + run % ix.K);
}

template <typename PT>
__global__ void __launch_bounds__(256)
range_nav_kernel(const FlatDev ix, int op, const u64* __restrict__ lo, const u64* __restrict__ hi,
                 const uint8_t* __restrict__ sym, u64 N, const u64* __restrict__ off, u64* __restrict__ out_a,
                 u64* __restrict__ out_b) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const u64 l = lo[t], r = hi[t];
    const uint8_t c = sym[t];
    if (op == RANGE_CLOSEST) {
        u64 res = ~0ull;
        if (l < ix.n) {
            u64 run = nav_run_of<PT>(ix, l);
            if (nav_head(ix, run) == c) {
                res = (u64)ld_pos<PT>(ix.start, run + 1) - 1;           // case 1: the range begins inside a c-run
            } else {
                for (++run; run < ix.r; ++run)                          // case 2: the first c-run after l
                    if (nav_head(ix, run) == c) { res = ld_pos<PT>(ix.start, run); break; }
            }
        }
        out_a[t] = res;
        return;
    }
    u64 cnt = 0;
    u64* oa = op == RANGE_BREAK_FILL ? out_a + off[t] : nullptr;
    u64* ob = op == RANGE_BREAK_FILL ? out_b + off[t] : nullptr;
    if (l <= r && r < ix.n) {
        const u64 rl = nav_run_of<PT>(ix, l), rr = nav_run_of<PT>(ix, r);
        if (nav_head(ix, rl) == c && nav_head(ix, rr) == c) {
            if (rl == rr) {
                if (oa) { oa[0] = l; ob[0] = r; }
                cnt = 1;
            } else {
                if (oa) { oa[0] = l; ob[0] = (u64)ld_pos<PT>(ix.start, rl + 1) - 1; }
                cnt = 1;
                for (u64 j = rl + 1; j < rr; ++j)
                    if (nav_head(ix, j) == c) {
                        if (oa) { oa[cnt] = ld_pos<PT>(ix.start, j); ob[cnt] = (u64)ld_pos<PT>(ix.start, j + 1) - 1; }
                        ++cnt;
                    }
                if (oa) { oa[cnt] = ld_pos<PT>(ix.start, rr); ob[cnt] = r; }
                ++cnt;
            }
        }
    }
    if (op == RANGE_BREAK_COUNT) out_a[t] = cnt;
}

// ---- multi-GPU fan-out helpers (contiguous shards re-cut at equal occurrence mass) --------------------
// n_occ(p) = hi - lo + 1, or 0 for the empty range {1,0} (r_index.hpp:307-313)
__global__ void __launch_bounds__(256) counts_kernel(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 N, u64* __restrict__ nocc) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) { const u64 l = lo[i], h = hi[i]; nocc[i] = h >= l ? h - l + 1 : 0; }
}

// Cut points of `shards` contiguous shards of near-equal WORK, work(p) = n_occ(p) + cost. One CTA: block-wide running
// prefix sum over the batch; where the prefix crosses k * total / shards the pattern that crosses goes to the side that
// leaves the smaller excess. Pure integer arithmetic — the same rule, bit for bit, as r-index_b200/_shard.py
// (balanced_cuts) and host/cli_common.hpp (GpuFleet::balanced_cuts):
//   i = the first pattern with cum(i) * shards >= total * k;  c = i + 1;
//   if (cum(i) * shards - total * k) > (total * k - cum(i - 1) * shards): c = i      (cum(-1) = 0)
//   cuts[k] = clamp(c, cuts[k-1], N)        (applied on the host: cuts must ascend)
// cuts_raw[k-1] = c for k = 1 .. shards-1. total * shards must stay below 2^64 (checked by the host: N * 2^40 * shards).
__global__ void __launch_bounds__(1024) balanced_cuts_kernel(const u64* __restrict__ nocc, u64 N, u32 shards, u64 cost,
                                                            u64* __restrict__ cuts_raw) {
    // every thread owns a contiguous chunk: chunk sums -> block-wide exclusive scan -> each thread walks its chunk
    // again looking for the crossings (two sweeps over 8 * N bytes by one CTA: ~10 us for 1e5 patterns)
    __shared__ u64 wsum[32];
    __shared__ u64 s_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u64 per = (N + blockDim.x - 1) / blockDim.x;
    const u64 i0 = min(N, (u64)threadIdx.x * per), i1 = min(N, i0 + per);
    if (threadIdx.x < shards - 1) cuts_raw[threadIdx.x] = ~0ull;   // no pattern crosses this target (shards <= 1024)
    u64 mine = 0;
    for (u64 i = i0; i < i1; ++i) mine += nocc[i] + cost;
    u64 inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u64 t = __shfl_up_sync(RIG_FULL, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        u64 ws = wsum[lane], wi = ws;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u64 t = __shfl_up_sync(RIG_FULL, wi, off);
            if (lane >= off) wi += t;
        }
        wsum[lane] = wi - ws;   // exclusive prefix of the warp sums
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    const u64 total = s_total;
    u64 cum = wsum[w] + inc - mine;   // work before this thread's chunk
    for (u64 i = i0; i < i1; ++i) {
        const u64 prev = cum;
        cum += nocc[i] + cost;
        for (u32 k = 1; k < shards; ++k) {      // does pattern i cross target k ?  prev * shards < total * k <= cum * shards
            const u64 t = total * k;
            if (prev * shards < t && cum * shards >= t) {
                u64 c = i + 1;
                if (cum * shards - t > t - prev * shards) c = i;
                cuts_raw[k - 1] = c;
            }
        }
    }
}

// The same cut points from the EXCLUSIVE PREFIX of the occurrence counts that the locate-mode search leaves behind
// (occ_off[0..N], occ_off[N] = total): cum(i) = occ_off[i + 1] + cost * (i + 1) is monotone, so each target is a binary
// search — no pass over the batch. out[0..S]: the clamped cuts; out[S+1..2S+1], out[2S+2..3S+2]: occ_off and ch_off at
// the cuts (what the host needs to size a shard's expansion). One warp; thread k searches target k.
__global__ void __launch_bounds__(32) cuts_from_offsets_kernel(const u64* __restrict__ occ_off, const u64* __restrict__ ch_off, u64 N,
                                                               u32 shards, u64 cost, u64* __restrict__ out) {
    __shared__ u64 s_c[1025];
    const u64 total = N ? occ_off[N] + cost * N : 0;
    for (u32 k = threadIdx.x; k <= shards; k += 32) {
        u64 c = N;
        if (k == 0) c = 0;
        else if (k < shards && N && total) {
            const u64 t = total * k;
            u64 lo = 0, hi = N;      // first i in [0, N) with cum(i) * shards >= t (cum(N-1) * shards = total * shards >= t)
            while (lo < hi) {
                const u64 mid = (lo + hi) >> 1;
                if ((occ_off[mid + 1] + cost * (mid + 1)) * shards >= t) hi = mid; else lo = mid + 1;
            }
            const u64 i = lo;
            if (i < N) {
                const u64 cum = occ_off[i + 1] + cost * (i + 1), prev = occ_off[i] + cost * i;
                c = (cum * shards - t > t - prev * shards) ? i : i + 1;
            }
        }
        s_c[k] = c;
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        u64 last = 0;
        for (u32 k = 0; k <= shards; ++k) {     // cuts must ascend
            u64 c = s_c[k];
            if (c < last) c = last;
            if (c > N) c = N;
            if (k == shards) c = N;
            last = c;
            out[k] = c;
            out[shards + 1 + k] = occ_off[c];
            out[2 * (shards + 1) + k] = ch_off[c];
        }
    }
}

// Counter block of a shard's expansion (phi_kernels.cuh: RIG_CTR_*): totals and bases from the batch-wide arrays.
__global__ void shard_setup_kernel(u64* __restrict__ ctr, const u64* __restrict__ occ_off, const u64* __restrict__ ch_off, u64 c0, u64 c1) {
    if (threadIdx.x < 16) ctr[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        ctr[2] = occ_off[c1] - occ_off[c0];   // RIG_CTR_TOTAL
        ctr[3] = ch_off[c1] - ch_off[c0];     // RIG_CTR_CHAINS
        ctr[9] = occ_off[c0];                 // RIG_CTR_OCC_BASE
        ctr[10] = ch_off[c0];                 // RIG_CTR_CH_BASE
    }
}

}  // namespace rigk
