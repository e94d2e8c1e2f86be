// search_kernels.cuh — sm_100a device code of the count + locate hot path (backward search, scans,
// digest). The Phi expansion kernel lives in phi_kernels.cuh.
//
//   search_kernel<G,LOCATE>  backward search, one group PAIR per pattern: G lanes evaluate
//                            rank(lo) and G lanes evaluate rank(hi+1) of the same LF step in
//                            lockstep (32/(2G) patterns per warp). LOCATE also tracks the toehold.
//                            Replaces r_index::count / count_and_get_occ / LF and
//                            rle_string::rank / operator[] / select / run_of_position
//                            (reference internal/r_index.hpp:171-190,292-302,482-545;
//                             internal/rle_string.hpp:126-256).
//   scan_*                   exclusive scans of n_occ and chain counts (output offsets).
//   digest_kernel            checksum of a u64 array.
//
// No tensor cores: the path has no dense contraction; it is dependent integer loads.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rigk {

typedef unsigned long long u64;
typedef unsigned int u32;

// Phi^1..Phi^D as one refined piecewise translation (flat_layout.hpp: PhiTable). Entries are RW words
// of u32 (FlatDev::w32, n < 2^32-1) or u64; RW = 4 (D=1), 8 (D=2,4), 16 (D=8).
struct PhiTabDev {
    const void* rec;    // [nbkt]   bucket records: D deltas, s1, nxt, cnt
    const void* pent;   // [pieces] piece entries:  D deltas, start
    u32 shift, D;
    u32 esz, packed;    // entry size in bytes; packed = 1: 64-bit index, D = 4, entries packed into 32 bytes
                        // (4 x 40-bit deltas | 40-bit s1 or start | 32-bit nxt | 24-bit cnt) instead of 8 x 8 bytes
};

// Phi^J as one piecewise translation (flat_layout.hpp: JumpTable), the seed table of the two-pass
// expansion. 8-word bucket records (delta0, s1, delta1, s2, nxt, cnt, 0, 0), 2-word piece entries.
struct SeedTabDev {
    const void* rec;
    const void* pent;
    u32 shift, J;       // J = 0: no seed table (single-pass expansion)
};

struct FlatDev {
    u64 n, r, nblk, toe0;
    u32 K, S, lf_shift, pad0;
    const u64* F;             // [257]
    const uint16_t* sid;      // [256]
    const void* start;        // [nblk*K+1]  PT (u32 when w32, else u64)
    const char* blk;          // [nblk] interleaved block records: K starts (PT), K heads (u8), S counts (PT)
    const void* last;         // [nblk*S]    PT: last run of each symbol before the block
    u32 blk_stride, off_head, off_cum, rec_w;  // rec_w: bytes per position word inside the records (4, 8, or 5 = 40-bit packed)
    const void* bstart;       // [nblk+1]    PT
    const u32* bdir;          // [lf_nbkt+1]
    const void* samples_last; // [r]         PT
    PhiTabDev phi;
    SeedTabDev seed;
    u32 w32, pad;             // w32: every position-holding array (and the Phi tables) uses 32-bit words; pad: diagnostics (0 in production)
    u64* dbg;                 // diagnostics: a 32 MB scratch window (RIG_VARIANT bit 15), else null
};

#define RIG_FULL 0xffffffffu

template <int G>
__device__ __forceinline__ u32 gballot(bool p, u32 gbase) {
    u32 m = __ballot_sync(RIG_FULL, p);
    return (G == 32) ? m : ((m >> gbase) & ((1u << G) - 1u));
}

template <int G, typename T>
__device__ __forceinline__ T greduce_add(T v) {
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(RIG_FULL, v, off);
    return v;
}

// Position-typed access to the flat arrays: PT = u32 when n < 2^32-1 (all arrays holding positions
// are then stored as 32-bit words: half the bytes per block, 32-bit compares/adds), u64 otherwise.
template <typename PT>
__device__ __forceinline__ PT ld_pos(const void* base, u64 idx) { return __ldg(reinterpret_cast<const PT*>(base) + idx); }

// Position word g of the record field that starts at byte `off` of record rp: 32-bit words, 64-bit words, or 40-bit
// packed little-endian words (FlatDev::rec_w = 5: two aligned 32-bit loads and a shift; the record is padded so
// that the second load stays inside it).
template <typename PT>
__device__ __forceinline__ PT rec_word(const FlatDev& ix, const char* rp, u32 off, u32 g) {
    if constexpr (sizeof(PT) == 4) {
        return __ldg(reinterpret_cast<const u32*>(rp + off) + g);
    } else {
        if (ix.rec_w == 8) return __ldg(reinterpret_cast<const u64*>(rp + off) + g);
        const u32 o = off + 5u * g;
        const u32* w = reinterpret_cast<const u32*>(rp + (o & ~3u));
        const u64 v = ((u64)__ldg(w + 1) << 32) | (u64)__ldg(w);
        return (v >> ((o & 3u) * 8u)) & 0xFFFFFFFFFFull;
    }
}

// One cooperative query by a group of G lanes (all 32 lanes of the warp execute this together,
// each group with its own x / c): locate the run holding BWT position x (0 <= x < n) and return
//   cnt        = #c in bwt[0..x]  (INCLUSIVE)   = rle_string::rank(x+1, c)   rle_string.hpp:170-218
//   run        = run holding x                   = rle_string::run_of_position rle_string.hpp:223-256
//   head_is_c  = bwt[x] == c                     = rle_string::operator[]     rle_string.hpp:126-131
//   prev_c_run = last run with head c strictly before `run` (or ~0): the run holding the last c
//                before x when bwt[x] != c — what rank/select/run_of_position compute together at
//                r_index.hpp:516-531.
// Dependent memory rounds: bdir -> (bstart, only when the bucket straddles blocks) -> block.
template <int G, bool WANT_RUN, typename PT>
__device__ __forceinline__ void block_query(const FlatDev& ix, PT x, uint8_t c, u32 sidc, int gl, u32 gbase,
                                            PT& cnt, u32& run, bool& head_is_c, u32& prev_c_run, bool want_run = true) {
    const u32 q = (u32)(x >> ix.lf_shift);
    u32 b0 = __ldg(ix.bdir + q);
    u32 b1 = __ldg(ix.bdir + q + 1);
    // G-ary search for the last block whose first position is <= x (blocks b0..b1 are candidates)
    while (__any_sync(RIG_FULL, b1 > b0)) {
        const u32 span = b1 - b0;
        const u32 step = (span + G - 1) / G;
        u32 probe = b0 + (u32)(gl + 1) * step;  // blocks < 2^32 / K: no overflow
        if (probe > b1 || probe < b0) probe = b1;
        const bool le = ld_pos<PT>(ix.bstart, probe) <= x;
        const u32 k = __popc(gballot<G>(le, gbase));
        if (span) {
            if (k == 0) {
                b1 = min(b1, b0 + step - 1);
            } else {
                const u32 nb0 = min(b1, b0 + k * step);
                b1 = min(b1, nb0 + step - 1);
                b0 = nb0;
            }
        }
    }
    const u32 base = b0 * G;
    const char* rp = ix.blk + (u64)b0 * ix.blk_stride;  // the block record: starts | heads | counts
    const PT st = rec_word<PT>(ix, rp, 0, (u32)gl);
    const uint8_t hd = __ldg(reinterpret_cast<const uint8_t*>(rp + ix.off_head) + gl);
    const PT cm = rec_word<PT>(ix, rp, ix.off_cum, sidc);
    const PT nxt = __shfl_down_sync(RIG_FULL, st, 1, G);
    const u32 mle = gballot<G>(st <= x, gbase);
    const int t = __popc(mle) - 1;  // >= 0: the block's first run starts at or before x
    const bool isc = (hd == c);
    PT contrib = 0;
    if (isc) contrib = (gl < t) ? (PT)(nxt - st) : ((gl == t) ? (PT)(x - st + 1) : (PT)0);
    // group-wide sum: log2(G) shuffle steps inside the group (a REDUX with a different lane mask per
    // group is serialised by the hardware, one pass per distinct mask — 8 passes per warp at G = 4)
    cnt = cm + greduce_add<G, PT>(contrib);
    if (WANT_RUN) {
        const u32 mc = gballot<G>(isc, gbase);
        head_is_c = (mc >> t) & 1u;
        const u32 below = mc & ((1u << t) - 1u);
        // no c-run before `run` inside this block: the per-block side table has the last one before it
        // (only consumed on a toehold miss, r_index.hpp:516-533)
        prev_c_run = below ? (base + (31 - __clz(below)))
                           : (head_is_c ? 0u : (u32)ld_pos<PT>(ix.last, (u64)b0 * ix.S + sidc));
        run = base + t;
    }
}

template <int G, bool LOCATE, typename PT>
__global__ void __launch_bounds__(256)
search_kernel(const FlatDev ix, const uint8_t* __restrict__ patt, u64 N, u64 m, u64* __restrict__ lo_out,
              u64* __restrict__ hi_out, u64* __restrict__ toe_out, u64* __restrict__ jl_out,
              u64* __restrict__ nch_out, u64* __restrict__ nocc_out, u64* __restrict__ lf_steps) {
    // per symbol: F[c], F[c+1], dense id — one shared-memory read per LF step
    struct SymEnt { PT f0, f1; u32 sid, pad; };
    __shared__ SymEnt sSym[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sSym[i].f0 = (PT)ix.F[i]; sSym[i].f1 = (PT)ix.F[i + 1];  // F[256] = n fits: n < 2^32-1 when PT = u32
        sSym[i].sid = ix.sid[i]; sSym[i].pad = 0;
    }
    __syncthreads();

    constexpr int PPW = 32 / (2 * G);  // patterns per warp
    const int lane = threadIdx.x & 31;
    const u64 warp = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int which = (lane / G) & 1;  // 0: rank before lo, 1: rank up to hi
    const int gl = lane % G;
    const u32 gbase = lane & ~(G - 1);
    const int pairbase = lane & ~(2 * G - 1);
    const u64 p = warp * PPW + lane / (2 * G);
    bool alive = p < N;
    const uint8_t* P = patt + (alive ? p : 0) * m;

    PT lo = 0, hi = (PT)(ix.n - 1);  // full_range, r_index.hpp:155-160
    PT k = (PT)ix.toe0;              // SA[n-1], r_index.hpp:489
    u32 steps = 0;
    for (u64 i = 0; i < m; ++i) {
        if (!__any_sync(RIG_FULL, alive)) break;  // r_index.hpp:297 (early exit on empty range)
        const uint8_t c = alive ? __ldg(P + (m - 1 - i)) : 0;
        const SymEnt se = sSym[c];
        const PT Fc = se.f0, Fc1 = se.f1;
        const bool act = alive && (Fc < Fc1);  // r_index.hpp:174 (absent symbol -> {1,0})
        const bool valid = act && (which || lo > 0);
        const PT x = valid ? (which ? hi : (PT)(lo - 1)) : (PT)0;
        PT cnt;
        u32 run = 0, prevc = 0;
        bool hic = false;
        block_query<G, LOCATE, PT>(ix, x, c, valid ? se.sid : 0, gl, gbase, cnt, run, hic, prevc);
        if (!valid) cnt = 0;
        const PT A = __shfl_sync(RIG_FULL, cnt, pairbase);      // rank(lo, c)      r_index.hpp:178
        const PT B = __shfl_sync(RIG_FULL, cnt, pairbase + G);  // rank(hi+1, c)    r_index.hpp:181
        if (LOCATE) {
            hic = __shfl_sync(RIG_FULL, (int)hic, pairbase + G);
            prevc = __shfl_sync(RIG_FULL, prevc, pairbase + G);
        }
        if (alive) {
            steps += act ? 1u : 0u;
            if (!act || B == A) {  // r_index.hpp:175,184
                lo = 1; hi = 0; alive = false;
            } else {
                if (LOCATE) {
                    if (hic) k -= 1;                                  // r_index.hpp:505-509
                    else k = ld_pos<PT>(ix.samples_last, prevc);      // r_index.hpp:516-533
                }
                lo = Fc + A;        // r_index.hpp:186
                hi = Fc + B - 1;    // r_index.hpp:188
            }
        }
    }
    const bool leader = (lane == pairbase) && (p < N);
    if (LOCATE) {
        // runs holding lo and hi: the range is cut into one Phi chain per overlapped run
        const bool ne = (p < N) && hi >= lo;
        PT cnt;
        u32 run = 0, prevc;
        bool hic;
        block_query<G, true, PT>(ix, ne ? (which ? hi : lo) : (PT)0, 0, 0, gl, gbase, cnt, run, hic, prevc);
        const u32 jL = __shfl_sync(RIG_FULL, run, pairbase);
        const u32 jR = __shfl_sync(RIG_FULL, run, pairbase + G);
        if (leader) {
            toe_out[p] = k;
            jl_out[p] = jL;
            nch_out[p] = ne ? (u64)(jR - jL + 1) : 0;
            nocc_out[p] = ne ? (u64)(hi - lo) + 1 : 0;  // r_index.hpp:338
        }
    }
    if (leader) { lo_out[p] = lo; hi_out[p] = hi; }
    // executed LF steps (for the algorithmic-bytes figure)
    u32 s = leader ? steps : 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(RIG_FULL, s, off);
    if (lane == 0 && s) atomicAdd(lf_steps, (u64)s);
}


// ------------------------------------------------------------------ single-pass offsets (fused into the search)
// The one-lane-per-pattern kernel also produces the exclusive prefix sums of n_occ and of the chain counts
// (= where every pattern's occurrences and chains start) with a decoupled look-back over its own CTAs, so a
// locate call needs no scan launches: tile t publishes its aggregate and adds up the aggregates of the tiles
// before it. Tiles are handed out by an atomic ticket, so every tile a CTA waits for has already started (it is
// resident or finished): no deadlock whatever the grid size.
//   ws[0]                       ticket counter
//   ws[2 + 5t]                  status of tile t: 0 not ready, 1 aggregate available, 2 inclusive prefix available
//   ws[3 + 5t], ws[4 + 5t]      aggregate of tile t (occurrences, chains)
//   ws[5 + 5t], ws[6 + 5t]      inclusive prefix of tiles 0..t
// The values are written first, then a device-scope fence, then the status word; a reader polls the status word
// with relaxed device-scope loads, fences, and reads the values (the status / value protocol of the single-pass
// scan with separate status and value arrays).
#define RIG_TILE_WORDS 5

__device__ __forceinline__ u64 ld_relaxed_gpu(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Called by every thread of a CTA (NW warps, one tile of 128 patterns) with its (n_occ, chains) contribution;
// ex_*: the exclusive prefix over everything before the thread; incl_*: the inclusive prefix at the end of this tile (the grand totals in the last tile).
//
// In a locate batch every CTA finishes its search at about the same time, so a look-back that waits for a
// predecessor's PREFIX turns into a chain of rounds (tile t learns its prefix ~t / window rounds after tile 0:
// measured +12..16 us on the 782 tiles of config C2 with 32- and 128-wide windows). Instead every tile publishes
// its AGGREGATE as soon as it has it, and sums the aggregates of the tiles before it in its GROUP of
// RIG_TILE_GROUP tiles directly — 128 threads, independent loads, no chain; only the last tile of a group publishes
// an inclusive prefix, which the tiles of the next group add: the chain has one link per 1024 tiles (131072
// patterns), and batches that large run in waves anyway.
#define RIG_TILE_GROUP 1024u
template <int NW>   // warps per CTA
__device__ __forceinline__ void tile_exclusive_scan(u64* ws, u32 tile, u64 a, u64 b, u64& ex_a, u64& ex_b,
                                                    u64& incl_a, u64& incl_b) {
    __shared__ u64 s_wa[NW], s_wb[NW], s_ra[NW], s_rb[NW];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u64 ia = a, ib = b;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u64 xa = __shfl_up_sync(RIG_FULL, ia, off), xb = __shfl_up_sync(RIG_FULL, ib, off);
        if (lane >= off) { ia += xa; ib += xb; }
    }
    if (lane == 31) { s_wa[w] = ia; s_wb[w] = ib; }
    __syncthreads();
    u64 pa = 0, pb = 0, ta = 0, tb = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        if (j < w) { pa += s_wa[j]; pb += s_wb[j]; }
        ta += s_wa[j]; tb += s_wb[j];
    }
    u64* st = ws + 2;
    if (threadIdx.x == 0) {   // aggregate first: nobody waits for more than this tile's own search
        u64* me = st + (u64)tile * RIG_TILE_WORDS;
        st_relaxed_gpu(me + 1, ta); st_relaxed_gpu(me + 2, tb);
        __threadfence();
        st_relaxed_gpu(me, 1);
    }
    const u32 base = tile & ~(RIG_TILE_GROUP - 1u);   // first tile of this tile's group
    u64 va = 0, vb = 0;
    for (u32 t = base + threadIdx.x; t < tile; t += blockDim.x) {   // aggregates of the group's earlier tiles
        const u64* p = st + (u64)t * RIG_TILE_WORDS;
        while (ld_relaxed_gpu(p) == 0) {}
        __threadfence();
        va += ld_relaxed_gpu(p + 1); vb += ld_relaxed_gpu(p + 2);
    }
    if (base > 0 && threadIdx.x == blockDim.x - 1) {               // inclusive prefix of the groups before
        const u64* p = st + (u64)(base - 1) * RIG_TILE_WORDS;
        while (ld_relaxed_gpu(p) != 2) {}
        __threadfence();
        va += ld_relaxed_gpu(p + 3); vb += ld_relaxed_gpu(p + 4);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { va += __shfl_xor_sync(RIG_FULL, va, off); vb += __shfl_xor_sync(RIG_FULL, vb, off); }
    if (lane == 0) { s_ra[w] = va; s_rb[w] = vb; }
    __syncthreads();
    u64 sum_a = 0, sum_b = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) { sum_a += s_ra[j]; sum_b += s_rb[j]; }
    if (threadIdx.x == 0 && ((tile + 1u) & (RIG_TILE_GROUP - 1u)) == 0) {   // last tile of a group: prefix for the next group
        u64* me = st + (u64)tile * RIG_TILE_WORDS;
        st_relaxed_gpu(me + 3, sum_a + ta); st_relaxed_gpu(me + 4, sum_b + tb);
        __threadfence();
        st_relaxed_gpu(me, 2);
    }
    ex_a = sum_a + pa + ia - a;
    ex_b = sum_b + pb + ib - b;
    incl_a = sum_a + ta;
    incl_b = sum_b + tb;
}

// ------------------------------------------------------------------ one lane per pattern
// The cooperative kernel above was measured issue-bound at every group size tried (16 -> 8 -> 4 lanes per
// rank query: 0.50 -> 0.30 -> 0.19 ms on C2), each halving of the group paying off in full: the ballots,
// shuffles and reductions that make G lanes act as one are most of its ~180 instructions per LF step.
// With the K = 4 block records one lane can evaluate a rank query alone — 4 run starts in ONE 128/256-bit
// load, the 4 heads in one 32-bit load, the symbol's count in a third — in ~50 scalar instructions with no
// warp collective at all, so a warp carries 32 patterns instead of 4 and the kernel turns from issue-bound
// into latency-bound (2 dependent loads per LF step: directory, record). Both rank queries of an LF step
// are done by the same lane with their loads interleaved.
template <typename PT>
struct LaneRec {           // one K = 4 block record as a lane reads it (position words: PT)
    PT s0, s1, s2, s3;     // run starts (padding runs start at n)
    u32 heads;             // 4 run heads, one byte each
    PT before;             // #c in the BWT before the block
};

template <typename PT>
__device__ __forceinline__ u32 lane_block_of(const FlatDev& ix, PT x, u32 b0, u32 b1) {
    while (b0 < b1) {      // the bucket straddles blocks (clustered short runs: a few steps at most)
        const u32 mid = b0 + ((b1 - b0 + 1) >> 1);
        if (ld_pos<PT>(ix.bstart, mid) <= x) b0 = mid; else b1 = mid - 1;
    }
    return b0;
}

template <typename PT>
__device__ __forceinline__ void lane_load(const FlatDev& ix, u32 b, u32 sidc, LaneRec<PT>& r) {
    const char* rp = ix.blk + (u64)b * ix.blk_stride;
    if constexpr (sizeof(PT) == 4) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(rp));
        r.s0 = v.x; r.s1 = v.y; r.s2 = v.z; r.s3 = v.w;
    } else if (ix.rec_w == 8) {
        const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(rp)), c = __ldg(reinterpret_cast<const ulonglong2*>(rp) + 1);
        r.s0 = a.x; r.s1 = a.y; r.s2 = c.x; r.s3 = c.y;
    } else {  // 40-bit packed: starts = bytes 0..19, heads = bytes 20..23: one 128-bit and one 64-bit load
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(rp));
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(rp) + 2);
        const u64 m40 = 0xFFFFFFFFFFull;
        r.s0 = (((u64)v.y << 32) | v.x) & m40;
        r.s1 = ((((u64)v.z << 32) | v.y) >> 8) & m40;
        r.s2 = ((((u64)v.w << 32) | v.z) >> 16) & m40;
        r.s3 = ((((u64)t.x << 32) | v.w) >> 24) & m40;
        r.heads = t.y;
        r.before = rec_word<PT>(ix, rp, ix.off_cum, sidc);
        return;
    }
    r.heads = __ldg(reinterpret_cast<const u32*>(rp + ix.off_head));
    r.before = rec_word<PT>(ix, rp, ix.off_cum, sidc);
}

// rank / run / head of position x from its block record (same outputs as block_query); all arithmetic in PT
template <bool WANT_RUN, typename PT>
__device__ __forceinline__ void lane_eval(const FlatDev& ix, const LaneRec<PT>& r, u32 b, PT x, uint8_t c, u32 sidc,
                                          PT& cnt, u32& run, bool& head_is_c, u32& prev_c_run, bool want_run = true) {
    const u32 t = (u32)(r.s1 <= x) + (u32)(r.s2 <= x) + (u32)(r.s3 <= x);  // run of the block holding x
    const u32 cc = (u32)c * 0x01010101u, eq = r.heads ^ cc;                 // byte g is 0 iff head g == c
    const bool m0 = (eq & 0xffu) == 0, m1 = (eq & 0xff00u) == 0, m2 = (eq & 0xff0000u) == 0, m3 = (eq & 0xff000000u) == 0;
    PT add = 0;
    if (m0 && t > 0) add += r.s1 - r.s0;
    if (m1 && t > 1) add += r.s2 - r.s1;
    if (m2 && t > 2) add += r.s3 - r.s2;
    const PT st = t == 0 ? r.s0 : (t == 1 ? r.s1 : (t == 2 ? r.s2 : r.s3));
    const bool mt = t == 0 ? m0 : (t == 1 ? m1 : (t == 2 ? m2 : m3));
    if (mt) add += x - st + 1;
    cnt = r.before + add;
    if (WANT_RUN && want_run) {
        const u32 mc = (u32)m0 | ((u32)m1 << 1) | ((u32)m2 << 2) | ((u32)m3 << 3);
        head_is_c = mt;
        const u32 below = mc & ((1u << t) - 1u);
        prev_c_run = below ? (b * 4u + (31u - __clz(below)))
                           : (mt ? 0u : (u32)ld_pos<PT>(ix.last, (u64)b * ix.S + sidc));
        run = b * 4u + t;
    }
}

template <bool LOCATE, typename PT>
__global__ void __launch_bounds__(128)
search_lane_kernel(const FlatDev ix, const uint8_t* __restrict__ patt, u64 N, u64 m, u64* __restrict__ lo_out,
                   u64* __restrict__ hi_out, u64* __restrict__ toe_out, u64* __restrict__ jl_out,
                   u64* __restrict__ choff_out, u64* __restrict__ occoff_out, u64* __restrict__ lf_steps,
                   u64* __restrict__ tile_ws, u64* __restrict__ totals) {
    struct SymEnt { PT f0, f1; u32 sid, pad; };  // F[256] = n fits: n < 2^32-1 when PT = u32
    __shared__ SymEnt sSym[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sSym[i].f0 = (PT)ix.F[i]; sSym[i].f1 = (PT)ix.F[i + 1]; sSym[i].sid = ix.sid[i]; sSym[i].pad = 0;
    }
    // locate: tiles (= 128 patterns) are taken by ticket, so that the look-back of the fused offset scan only
    // waits for CTAs that have started (tile_exclusive_scan)
    __shared__ u32 s_tile;
    if (LOCATE && threadIdx.x == 0) s_tile = (u32)atomicAdd(tile_ws, 1ull);
    __syncthreads();
    const u32 tile = LOCATE ? s_tile : blockIdx.x;
    const u64 p = (u64)tile * blockDim.x + threadIdx.x;
    bool alive = p < N;
    const uint8_t* P = patt + (alive ? p : 0) * m;
    PT lo = 0, hi = (PT)(ix.n - 1);  // full_range, r_index.hpp:155-160
    PT k = (PT)ix.toe0;              // SA[n-1], r_index.hpp:489
    u32 steps = 0;
    for (u64 i = 0; i < m; ++i) {
        if (!__any_sync(RIG_FULL, alive)) break;  // r_index.hpp:297 (early exit on empty range)
        if (!alive) continue;
        const uint8_t c = __ldg(P + (m - 1 - i));
        const SymEnt se = sSym[c];
        if (!(se.f0 < se.f1)) { lo = 1; hi = 0; alive = false; continue; }  // r_index.hpp:174 (absent symbol -> {1,0})
        ++steps;
        // rank(lo, c) = #c in bwt[0, lo) and rank(hi + 1, c): r_index.hpp:178,181 — both queries' loads interleaved
        const bool need_a = lo > 0;
        const PT xa = need_a ? (PT)(lo - 1) : (PT)0, xb = hi;
        const u32 qa = (u32)(xa >> ix.lf_shift), qb = (u32)(xb >> ix.lf_shift);
        const u32 a0 = __ldg(ix.bdir + qa), a1 = __ldg(ix.bdir + qa + 1), b0 = __ldg(ix.bdir + qb), b1 = __ldg(ix.bdir + qb + 1);
        const u32 ba = lane_block_of<PT>(ix, xa, a0, a1), bb = lane_block_of<PT>(ix, xb, b0, b1);
        LaneRec<PT> ra, rb;
        lane_load<PT>(ix, ba, se.sid, ra);
        lane_load<PT>(ix, bb, se.sid, rb);
        PT A = 0, B;
        u32 run = 0, prevc = 0;
        bool hic = false;
        {
            u32 r0, p0; bool h0;
            lane_eval<false, PT>(ix, ra, ba, xa, c, se.sid, A, r0, h0, p0);
            if (!need_a) A = 0;
        }
        lane_eval<LOCATE, PT>(ix, rb, bb, xb, c, se.sid, B, run, hic, prevc);
        if (B == A) { lo = 1; hi = 0; alive = false; continue; }  // r_index.hpp:175,184
        if (LOCATE) {
            if (hic) k -= 1;                                   // r_index.hpp:505-509
            else k = ld_pos<PT>(ix.samples_last, prevc);       // r_index.hpp:516-533
        }
        lo = se.f0 + A;        // r_index.hpp:186
        hi = se.f0 + B - 1;    // r_index.hpp:188
    }
    if (LOCATE) {  // runs holding lo and hi: the range is cut into one Phi chain per overlapped run
        const bool ne = (p < N) && hi >= lo;
        u32 jL = 0, jR = 0;
        if (ne) {
            const u32 qa = (u32)(lo >> ix.lf_shift), qb = (u32)(hi >> ix.lf_shift);
            const u32 ba = lane_block_of<PT>(ix, lo, __ldg(ix.bdir + qa), __ldg(ix.bdir + qa + 1));
            const u32 bb = lane_block_of<PT>(ix, hi, __ldg(ix.bdir + qb), __ldg(ix.bdir + qb + 1));
            LaneRec<PT> ra, rb;
            lane_load<PT>(ix, ba, 0, ra);
            lane_load<PT>(ix, bb, 0, rb);
            jL = ba * 4u + (u32)(ra.s1 <= lo) + (u32)(ra.s2 <= lo) + (u32)(ra.s3 <= lo);
            jR = bb * 4u + (u32)(rb.s1 <= hi) + (u32)(rb.s2 <= hi) + (u32)(rb.s3 <= hi);
        }
        const u64 nocc = ne ? (u64)(hi - lo) + 1 : 0;   // r_index.hpp:338
        const u64 nch = ne ? (u64)(jR - jL + 1) : 0;
        u64 ex_occ, ex_ch, in_occ, in_ch;
        tile_exclusive_scan<4>(tile_ws, tile, nocc, nch, ex_occ, ex_ch, in_occ, in_ch);
        if (p < N) {
            toe_out[p] = (u64)k;
            jl_out[p] = jL;
            occoff_out[p] = ex_occ;
            choff_out[p] = ex_ch;
        }
        if ((u64)(tile + 1) * blockDim.x >= N && threadIdx.x == 0) {  // the last tile closes both arrays
            occoff_out[N] = in_occ; choff_out[N] = in_ch;
            totals[0] = in_occ; totals[1] = in_ch;
        }
    }
    if (p < N) { lo_out[p] = (u64)lo; hi_out[p] = (u64)hi; }
    // executed LF steps (for the algorithmic-bytes figure)
    u32 sum = steps;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(RIG_FULL, sum, off);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(lf_steps, (u64)sum);
}

// ------------------------------------------------------------------ two lanes per pattern
// The lane kernel is latency-bound with the machine half empty on batches of ~1e5 patterns (config C2: 782 CTAs of
// 4 warps, 32% occupancy, ~300 warp instructions per LF step at ~22 cycles each). Here the two rank queries of an
// LF step sit on ADJACENT lanes — even lane: rank(lo), odd lane: rank(hi + 1) and the toehold — and meet in one
// shuffle: half the instructions per step on every lane's critical path, twice the warps in flight. Both lanes of a
// pair keep identical (lo, hi, alive), so the pair leaves the loop together.
template <bool LOCATE, typename PT>
__global__ void __launch_bounds__(256)
search_pair_kernel(const FlatDev ix, const uint8_t* __restrict__ patt, u64 N, u64 m, u64* __restrict__ lo_out,
                   u64* __restrict__ hi_out, u64* __restrict__ toe_out, u64* __restrict__ jl_out,
                   u64* __restrict__ choff_out, u64* __restrict__ occoff_out, u64* __restrict__ lf_steps,
                   u64* __restrict__ tile_ws, u64* __restrict__ totals) {
    struct SymEnt { PT f0, f1; u32 sid, pad; };
    __shared__ SymEnt sSym[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sSym[i].f0 = (PT)ix.F[i]; sSym[i].f1 = (PT)ix.F[i + 1]; sSym[i].sid = ix.sid[i]; sSym[i].pad = 0;
    }
    __shared__ u32 s_tile;
    if (LOCATE && threadIdx.x == 0) s_tile = (u32)atomicAdd(tile_ws, 1ull);   // tiles by ticket: see search_lane_kernel
    __syncthreads();
    const u32 tile = LOCATE ? s_tile : blockIdx.x;
    const u32 role = threadIdx.x & 1u;
    const u64 p = (u64)tile * 128u + (threadIdx.x >> 1);
    bool alive = p < N;
    const uint8_t* P = patt + (alive ? p : 0) * m;
    PT lo = 0, hi = (PT)(ix.n - 1);  // full_range, r_index.hpp:155-160
    PT k = (PT)ix.toe0;              // SA[n-1], r_index.hpp:489 (kept by the odd lane)
    u32 steps = 0;
    for (u64 i = 0; i < m; ++i) {
        const u32 amask = __ballot_sync(RIG_FULL, alive);
        if (!amask) break;           // r_index.hpp:297 (early exit on empty range)
        if (!alive) continue;
        const uint8_t c = __ldg(P + (m - 1 - i));
        const SymEnt se = sSym[c];
        const bool sym = se.f0 < se.f1;            // absent symbol: both counts 0 -> {1,0} below (r_index.hpp:174)
        const u32 sidc = sym ? se.sid : 0u;
        steps += (u32)(sym && !role);
        // even lane: #c in bwt[0, lo); odd lane: #c in bwt[0, hi]  (r_index.hpp:178,181)
        const bool need = role || lo > 0;
        const PT x = role ? hi : (lo > 0 ? (PT)(lo - 1) : (PT)0);
        const u32 q = (u32)(x >> ix.lf_shift);
        const u32 b = lane_block_of<PT>(ix, x, __ldg(ix.bdir + q), __ldg(ix.bdir + q + 1));
        LaneRec<PT> r;
        lane_load<PT>(ix, b, sidc, r);
        PT cnt = 0;
        u32 run = 0, prevc = 0;
        bool hic = false;
        lane_eval<LOCATE, PT>(ix, r, b, x, c, sidc, cnt, run, hic, prevc, role != 0);
        if (!need || !sym) cnt = 0;
        const PT other = __shfl_xor_sync(amask, cnt, 1);
        const PT A = role ? other : cnt, B = role ? cnt : other;
        if (B == A) { lo = 1; hi = 0; alive = false; continue; }  // r_index.hpp:175,184
        if (LOCATE && role) {
            if (hic) k -= 1;                                   // r_index.hpp:505-509
            else k = ld_pos<PT>(ix.samples_last, prevc);       // r_index.hpp:516-533
        }
        lo = se.f0 + A;        // r_index.hpp:186
        hi = se.f0 + B - 1;    // r_index.hpp:188
    }
    if (LOCATE) {  // even lane: the run holding lo; odd lane: the run holding hi — one Phi chain per overlapped run
        const bool ne = (p < N) && hi >= lo;
        u32 j = 0;
        if (ne) {
            const PT x = role ? hi : lo;
            const u32 q = (u32)(x >> ix.lf_shift);
            const u32 b = lane_block_of<PT>(ix, x, __ldg(ix.bdir + q), __ldg(ix.bdir + q + 1));
            LaneRec<PT> r;
            lane_load<PT>(ix, b, 0, r);
            j = b * 4u + (u32)(r.s1 <= x) + (u32)(r.s2 <= x) + (u32)(r.s3 <= x);
        }
        const u32 jo = __shfl_xor_sync(RIG_FULL, j, 1);
        const u32 jL = role ? jo : j, jR = role ? j : jo;
        const u64 nocc = (ne && !role) ? (u64)(hi - lo) + 1 : 0;   // r_index.hpp:338
        const u64 nch = (ne && !role) ? (u64)(jR - jL + 1) : 0;
        u64 ex_occ, ex_ch, in_occ, in_ch;
        tile_exclusive_scan<8>(tile_ws, tile, nocc, nch, ex_occ, ex_ch, in_occ, in_ch);
        if (p < N) {
            if (role) toe_out[p] = (u64)k;
            else { jl_out[p] = jL; occoff_out[p] = ex_occ; choff_out[p] = ex_ch; }
        }
        if ((u64)(tile + 1) * 128u >= N && threadIdx.x == 0) {  // the last tile closes both arrays
            occoff_out[N] = in_occ; choff_out[N] = in_ch;
            totals[0] = in_occ; totals[1] = in_ch;
        }
    }
    if (p < N) { if (role) hi_out[p] = (u64)hi; else lo_out[p] = (u64)lo; }
    u32 sum = steps;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(RIG_FULL, sum, off);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(lf_steps, (u64)sum);
}

// ------------------------------------------------------------------ scans
#define RIG_SCAN_THREADS 256
#define RIG_SCAN_ITEMS 8
#define RIG_SCAN_TILE (RIG_SCAN_THREADS * RIG_SCAN_ITEMS)

__device__ __forceinline__ u64 block_reduce_u64(u64 v, u64* smem) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(RIG_FULL, v, off);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) smem[w] = v;
    __syncthreads();
    u64 t = 0;
    if (w == 0) {
        t = (l < (blockDim.x >> 5)) ? smem[l] : 0;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) t += __shfl_xor_sync(RIG_FULL, t, off);
    }
    __syncthreads();
    return t;  // valid in warp 0
}

// tile sums of two arrays: sums[0][b], sums[1][b]
__global__ void __launch_bounds__(RIG_SCAN_THREADS)
scan_tile_sums(const u64* __restrict__ a, const u64* __restrict__ b, u64 N, u64* __restrict__ sums, u64 ntiles) {
    __shared__ u64 sm[32];
    const u64 base = (u64)blockIdx.x * RIG_SCAN_TILE;
    u64 sa = 0, sb = 0;
    for (int i = 0; i < RIG_SCAN_ITEMS; ++i) {
        const u64 idx = base + (u64)i * RIG_SCAN_THREADS + threadIdx.x;
        if (idx < N) { sa += a[idx]; sb += b[idx]; }
    }
    sa = block_reduce_u64(sa, sm);
    sb = block_reduce_u64(sb, sm);
    if (threadIdx.x == 0) { sums[blockIdx.x] = sa; sums[ntiles + blockIdx.x] = sb; }
}

// single block: exclusive scan of the tile sums in place; totals[0], totals[1]
__global__ void __launch_bounds__(1024) scan_sums_inplace(u64* __restrict__ sums, u64 ntiles, u64* __restrict__ totals) {
    __shared__ u64 wsum[32];
    __shared__ u64 carry;
    for (int arr = 0; arr < 2; ++arr) {
        u64* s = sums + (u64)arr * ntiles;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (u64 base = 0; base < ntiles; base += blockDim.x) {
            const u64 idx = base + threadIdx.x;
            const u64 v = idx < ntiles ? s[idx] : 0;
            u64 inc = v;
            const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const u64 t = __shfl_up_sync(RIG_FULL, inc, off);
                if (l >= off) inc += t;
            }
            if (l == 31) wsum[w] = inc;
            __syncthreads();
            if (w == 0) {
                u64 ws = wsum[l];
                u64 wi = ws;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const u64 t = __shfl_up_sync(RIG_FULL, wi, off);
                    if (l >= off) wi += t;
                }
                wsum[l] = wi - ws;  // exclusive prefix of warp sums
            }
            __syncthreads();
            const u64 c0 = carry;
            const u64 excl = c0 + wsum[w] + inc - v;
            if (idx < ntiles) s[idx] = excl;
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) carry = excl + v;
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[arr] = carry;
        __syncthreads();
    }
}

// per tile: exclusive scan with the tile's offset; out arrays have N+1 entries
__global__ void __launch_bounds__(RIG_SCAN_THREADS)
scan_tiles(const u64* __restrict__ a, const u64* __restrict__ b, u64 N, const u64* __restrict__ sums, u64 ntiles,
           u64* __restrict__ out_a, u64* __restrict__ out_b, const u64* __restrict__ totals,
           const u64* __restrict__ base_a) {  // *base_a (if given) is added to every out_a entry
    __shared__ u64 wsum[2][RIG_SCAN_THREADS / 32];
    const u64 base = (u64)blockIdx.x * RIG_SCAN_TILE + (u64)threadIdx.x * RIG_SCAN_ITEMS;
    u64 va[RIG_SCAN_ITEMS], vb[RIG_SCAN_ITEMS];
    u64 ta = 0, tb = 0;
#pragma unroll
    for (int i = 0; i < RIG_SCAN_ITEMS; ++i) {
        const u64 idx = base + i;
        va[i] = idx < N ? a[idx] : 0;
        vb[i] = idx < N ? b[idx] : 0;
        ta += va[i]; tb += vb[i];
    }
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    u64 ia = ta, ib = tb;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u64 xa = __shfl_up_sync(RIG_FULL, ia, off), xb = __shfl_up_sync(RIG_FULL, ib, off);
        if (l >= off) { ia += xa; ib += xb; }
    }
    if (l == 31) { wsum[0][w] = ia; wsum[1][w] = ib; }
    __syncthreads();
    u64 wa = 0, wb = 0;
    for (int j = 0; j < w; ++j) { wa += wsum[0][j]; wb += wsum[1][j]; }
    const u64 ba = base_a ? __ldcg(base_a) : 0;
    u64 ea = ba + sums[blockIdx.x] + wa + ia - ta;
    u64 eb = sums[ntiles + blockIdx.x] + wb + ib - tb;
#pragma unroll
    for (int i = 0; i < RIG_SCAN_ITEMS; ++i) {
        const u64 idx = base + i;
        if (idx < N) { out_a[idx] = ea; out_b[idx] = eb; }
        ea += va[i]; eb += vb[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out_a[N] = ba + totals[0]; out_b[N] = totals[1]; }
}

// 64-bit positions -> 32-bit positions (rig_locate_batch32: every position < n < 2^32), 4 per thread
__global__ void __launch_bounds__(256) narrow_kernel(const u64* __restrict__ in, u32* __restrict__ out, u64 count) {
    const u64 i = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= count) {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(in + i), b = *reinterpret_cast<const ulonglong2*>(in + i + 2);
        *reinterpret_cast<uint4*>(out + i) = make_uint4((u32)a.x, (u32)a.y, (u32)b.x, (u32)b.y);
    } else {
        for (u64 k = i; k < count; ++k) out[k] = (u32)in[k];
    }
}

// out[0] += sum v, out[1] += sum v*(i+1)   (mod 2^64)
__global__ void __launch_bounds__(256) digest_kernel(const u64* __restrict__ v, u64 count, u64* __restrict__ out) {
    __shared__ u64 sm[32];
    u64 s0 = 0, s1 = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x) {
        const u64 x = v[i];
        s0 += x; s1 += x * (i + 1);
    }
    s0 = block_reduce_u64(s0, sm);
    s1 = block_reduce_u64(s1, sm);
    if (threadIdx.x == 0) { atomicAdd(out, s0); atomicAdd(out + 1, s1); }
}

}  // namespace rigk
