// flat_layout.hpp — "flatten once at load": logical r-index arrays -> the word arrays the
// kernels read. Plain C++ (no CUDA), so the layout can also be checked on a CPU-only box by
// tests/support/flat_check.cpp (a test double that walks these same arrays with scalar code).
//
// What the reference keeps as SDSL objects, and what replaces each here (all O(r) words):
//
//   reference member (internal/…)                      flat arrays
//   -------------------------------------------------  ---------------------------------------------
//   rle_string::runs  (EF marks, every B=2nd run end)   start[]  : u64 start position of every run,
//     rle_string.hpp:78,112                               grouped in BLOCKS of K runs (K = lanes that
//   rle_string::runs_per_letter[256] (EF, per letter)     cooperate on one rank query); bstart[] = first
//     rle_string.hpp:115-117                              start of each block; bdir[] = direct-addressed
//   rle_string::run_heads (Huffman wavelet tree)          directory position>>s -> block (the role of
//     rle_string.hpp:119, huff_string.hpp                 sd_vector's high-bits select)
//                                                       head[]   : u8 run heads, same block order
//                                                       cum[]    : per block, per symbol: (#symbol in
//                                                         BWT before the block, id of the last run of
//                                                         that symbol before the block) — the rank
//                                                         DIRECTORY, interleaved at block granularity.
//                                                         One level replaces the ℓ wavelet-tree levels
//                                                         AND the per-letter Elias-Fano select.
//   r_index::pred (EF) + pred_to_run + samples_last      PhiTable: Phi is a piecewise translation of
//     r_index.hpp:663-665                                  [0,n): for i in (p_k, p_{k+1}] it adds the
//                                                         constant samples_last[run_k - 1] - p_k (mod n).
//                                                         Composing it with itself D times refines the
//                                                         pieces (<= D*r of them); each piece then holds
//                                                         D deltas, so ONE record lookup yields
//                                                         SA[x-1], ..., SA[x-D] from SA[x].
//                                                         Bucket records (deltas of the piece covering
//                                                         the bucket start, next piece start, next piece
//                                                         id) are direct-addressed by position>>s.
//   r_index::samples_last                               samples_last[] : u64, run order (toeholds,
//                                                         chain splitting at run boundaries)
//   r_index::F                                          F[257], sid[256] (symbol -> dense id)
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cstring>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../../include/rindex_gpu.h"

namespace rigf {

typedef uint64_t u64;
typedef uint32_t u32;

// Phi, Phi^2, ..., Phi^D as one refined piecewise translation.
//   piece k covers [start[k], start[k+1]);  Phi^j(i) = (i + delta[k*D + j-1]) mod n  for i in piece k.
// Device form: two arrays of RW-word entries (RW = 4 for D=1, 8 for D=2,4, 16 for D=8), so that a
// lane needs ONE load instruction per step whatever it is doing:
//   rec[q]  bucket q = positions [q<<shift, (q+1)<<shift):
//           [0..D) deltas of the piece covering the bucket's first position
//           [D]    s1  = start of the first piece that begins inside the bucket (after its first position), else ~0
//           [D+1]  nxt = index of that piece
//           [D+2]  cnt = number of pieces that begin inside the bucket
//   pent[k] piece k:  [0..D) its deltas, [D] its start
// query i: i < s1 -> record deltas; else the answer is the last piece in [nxt, nxt+cnt) with start <= i
// (cnt == 1: piece nxt; otherwise a binary search whose every probe is one pent[] load).
struct PhiTable {
    u32 D = 1, RW = 4, shift = 0;
    u64 nbkt = 0;
    std::vector<u64> start;   // [pieces], start[0] == 0, strictly ascending
    std::vector<u64> delta;   // [pieces * D]
    std::vector<u64> rec;     // [nbkt * RW]
    std::vector<u64> pent;    // [pieces * RW]
    bool packable = true;     // D = 6: every bucket's (nxt, cnt) fits the shared word
    u64 pieces() const { return start.size(); }
    u64 bytes(bool w32) const { return (rec.size() + pent.size()) * (w32 ? 4 : 8); }
    static u32 record_words(u32 D) { return D == 1 ? 4 : (D <= 6 ? 8 : 16); }
    // D = 6 ("six occurrences per 32-byte entry", 32-bit words only): the 6 deltas and s1 / start fill 7 of the 8 words,
    // so nxt and cnt share the last one: nxt in the low 24 bits, cnt in the high 8 (the flatten step falls back to
    // D = 4 when a table has 2^24 pieces or a bucket holds 255 piece starts).
    static u64 nxt_of(const u64* w, u32 D) { return D == 6 ? (w[7] & 0xFFFFFFull) : w[D + 1]; }
    static u64 cnt_of(const u64* w, u32 D) { return D == 6 ? ((w[7] >> 24) & 0xFFull) : w[D + 2]; }
    u64 piece_of(u64 i) const { return (u64)(std::upper_bound(start.begin(), start.end(), i) - start.begin()) - 1; }
    // scalar evaluation of Phi^j(i), 1 <= j <= D (host-side construction + tests)
    u64 apply(u64 i, u32 j, u64 n) const {
        u64 v = i + delta[piece_of(i) * D + (j - 1)];
        return v >= n ? v - n : v;
    }
    void build_directory(u64 n, u32 buckets_log2) {
        RW = record_words(D);
        shift = 0;
        u64 target = pieces() << buckets_log2;
        if (target < 1) target = 1;
        while (((n - 1) >> shift) + 1 > target) ++shift;
        nbkt = ((n - 1) >> shift) + 1;
        rec.assign(nbkt * RW, 0);
        const u64 P = pieces();
        pent.assign(P * RW, 0);
        unsigned T = std::thread::hardware_concurrency();
        if (T > 32) T = 32;
        if (T < 1 || nbkt < (1u << 16)) T = 1;
        packable = true;
        auto fill = [&](u64 q0, u64 q1, u64 k0, u64 k1) {   // (`packable` is only ever cleared: a benign race between the fill threads)
            for (u64 k = k0; k < k1; ++k) {
                for (u32 j = 0; j < D; ++j) pent[k * RW + j] = delta[k * D + j];
                pent[k * RW + D] = start[k];
            }
            if (q0 >= q1) return;
            u64 a = piece_of(q0 << shift);  // the piece covering the first bucket's first position
            for (u64 q = q0; q < q1; ++q) {
                const u64 lo = q << shift, hi = (q + 1) << shift;
                while (a + 1 < P && start[a + 1] <= lo) ++a;
                u64 e = a + 1;
                while (e < P && start[e] < hi) ++e;  // pieces a+1 .. e-1 begin inside the bucket
                u64* R = &rec[q * RW];
                for (u32 j = 0; j < D; ++j) R[j] = delta[a * D + j];
                R[D] = (e > a + 1) ? start[a + 1] : ~(u64)0;
                const u64 nxt = (e > a + 1) ? a + 1 : 0, cnt = e - (a + 1);
                if (D == 6) {
                    if (cnt >= 255 || nxt >= (1ull << 24)) packable = false;
                    R[7] = (nxt & 0xFFFFFFull) | ((cnt & 0xFFull) << 24);
                } else {
                    R[D + 1] = nxt;
                    R[D + 2] = cnt;
                }
            }
        };
        if (T == 1) { fill(0, nbkt, 0, P); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t]() { fill(nbkt * t / T, nbkt * (t + 1) / T, P * t / T, P * (t + 1) / T); });
        for (auto& x : th) x.join();
    }
};

// Extend a table holding Phi^1..Phi^j to Phi^1..Phi^(j+1): every piece is cut where its image under
// Phi^j crosses a piece boundary of Phi (the single-step table), and gets delta_{j+1} = delta_j + delta_Phi.
static inline void extend_range(const PhiTable& A, const PhiTable& Phi, u64 n, u64 k0, u64 k1, std::vector<u64>& Cs,
                                std::vector<u64>& Cd) {
    const u32 j = A.D;
    const u64 PA = A.pieces(), PB = Phi.pieces();
    Cs.reserve((k1 - k0) * 2 + 16); Cd.reserve(((k1 - k0) * 2 + 16) * (j + 1));
    for (u64 k = k0; k < k1; ++k) {
        const u64 s = A.start[k], e = (k + 1 < PA) ? A.start[k + 1] : n, a = A.delta[k * j + (j - 1)];
        const u64 len = e - s;
        u64 u = s + a;  // image of s under Phi^j
        if (u >= n) u -= n;
        u64 done = 0;
        while (done < len) {  // at most two linear ranges (the image may wrap around n)
            const bool wrapped = u + done >= n;
            const u64 pos = wrapped ? u + done - n : u + done;
            const u64 lin_end = wrapped ? (u + len - n) : std::min(u + len, n);
            u64 b = Phi.piece_of(pos);
            u64 cur = pos;
            while (cur < lin_end) {
                const u64 pend = std::min(lin_end, (b + 1 < PB) ? Phi.start[b + 1] : n);
                u64 d = a + Phi.delta[b];
                if (d >= n) d -= n;
                Cs.push_back(s + done + (cur - pos));
                for (u32 t = 0; t < j; ++t) Cd.push_back(A.delta[k * j + t]);
                Cd.push_back(d);
                cur = pend; ++b;
            }
            done += lin_end - pos;
        }
    }
}

// (pieces of A are independent: large tables are cut into contiguous chunks, one host thread each)
static inline PhiTable extend_by_phi(const PhiTable& A, const PhiTable& Phi, u64 n) {
    PhiTable C;
    C.D = A.D + 1;
    const u64 PA = A.pieces();
    unsigned T = std::thread::hardware_concurrency();
    if (T > 32) T = 32;
    if (T < 2 || PA < (1u << 16)) { extend_range(A, Phi, n, 0, PA, C.start, C.delta); return C; }
    std::vector<std::vector<u64>> ps(T), pd(T);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t]() { extend_range(A, Phi, n, PA * t / T, PA * (t + 1) / T, ps[t], pd[t]); });
    for (auto& x : th) x.join();
    u64 total = 0;
    for (unsigned t = 0; t < T; ++t) total += ps[t].size();
    C.start.reserve(total); C.delta.reserve(total * C.D);
    for (unsigned t = 0; t < T; ++t) {
        C.start.insert(C.start.end(), ps[t].begin(), ps[t].end());
        C.delta.insert(C.delta.end(), pd[t].begin(), pd[t].end());
        std::vector<u64>().swap(ps[t]); std::vector<u64>().swap(pd[t]);
    }
    return C;
}

// Phi^J as ONE piecewise translation (final delta only): the seed table of the two-pass expansion.
//   piece k covers [start[k], start[k+1]);  Phi^J(i) = (i + delta[k]) mod n  for i in piece k.
// Built by doubling (Phi^2 = Phi o Phi, Phi^4 = Phi^2 o Phi^2, ...): J is a power of two and the
// number of pieces is sum over runs of min(J, run length) <= min(J*r, n).
// Device form: one 16-word bucket record per direct-addressed bucket (one bucket per piece on average),
// resolving up to SEED_INLINE pieces that begin inside the bucket without a second load — a hop of the
// seed pass is a chain of dependent DRAM-latency loads, so the record is sized to make it ONE load —
// + a 2-word entry per piece for more crowded buckets:
//   rec[q]   [0] delta of the piece covering the bucket's first position
//            [1+2i] s_i = start of the i-th piece beginning inside the bucket (else ~0)   [2+2i] its delta   (i < 6)
//            [13] nxt = index of the piece starting at s_0   [14] cnt = pieces beginning inside   [15] 0
//   pent[k]  [0] delta  [1] start
// query v: the last inline piece with s_i <= v (else rec[0]); if cnt > 6 and v >= s_5, the last piece in
// [nxt+5, nxt+cnt) with start <= v (binary search over pent).
static const uint32_t SEED_INLINE = 6;
static const uint32_t SEED_RW = 16;
struct JumpTable {
    u32 J = 0, shift = 0;
    u64 nbkt = 0;
    std::vector<u64> start, delta;
    std::vector<u64> rec;   // [nbkt * SEED_RW]
    std::vector<u64> pent;  // [pieces * 2]
    u64 pieces() const { return start.size(); }
    u64 bytes(bool w32) const { return (rec.size() + pent.size()) * (w32 ? 4 : 8); }
    u64 piece_of(u64 i) const { return (u64)(std::upper_bound(start.begin(), start.end(), i) - start.begin()) - 1; }
    u64 apply(u64 i, u64 n) const {
        u64 v = i + delta[piece_of(i)];
        return v >= n ? v - n : v;
    }
    void build_directory(u64 n, u32 buckets_log2) {
        shift = 0;
        u64 target = pieces() << buckets_log2;
        if (target < 1) target = 1;
        while (((n - 1) >> shift) + 1 > target) ++shift;
        nbkt = ((n - 1) >> shift) + 1;
        const u64 P = pieces();
        rec.resize(nbkt * SEED_RW);
        pent.resize(P * 2);
        unsigned T = std::thread::hardware_concurrency();
        if (T > 32) T = 32;
        if (T < 1 || nbkt < (1u << 16)) T = 1;
        auto fill = [&](u64 q0, u64 q1, u64 k0, u64 k1) {
            for (u64 k = k0; k < k1; ++k) { pent[2 * k] = delta[k]; pent[2 * k + 1] = start[k]; }
            if (q0 >= q1) return;
            u64 a = piece_of(q0 << shift);  // the piece covering the first bucket's first position
            for (u64 q = q0; q < q1; ++q) {
                const u64 lo = q << shift, hi = (q + 1) << shift;
                while (a + 1 < P && start[a + 1] <= lo) ++a;
                u64 e = a + 1;
                while (e < P && start[e] < hi) ++e;  // pieces a+1 .. e-1 begin inside the bucket
                u64* R = &rec[q * SEED_RW];
                const u64 cnt = e - (a + 1);
                R[0] = delta[a];
                for (u32 i = 0; i < SEED_INLINE; ++i) {
                    R[1 + 2 * i] = i < cnt ? start[a + 1 + i] : ~(u64)0;
                    R[2 + 2 * i] = i < cnt ? delta[a + 1 + i] : 0;
                }
                R[13] = cnt >= 1 ? a + 1 : 0;
                R[14] = cnt;
                R[15] = 0;
            }
        };
        if (T == 1) { fill(0, nbkt, 0, P); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t)
            th.emplace_back([&, t]() { fill(nbkt * t / T, nbkt * (t + 1) / T, P * t / T, P * (t + 1) / T); });
        for (auto& x : th) x.join();
    }
};

// C = B o A (apply A, then B) for two piecewise translations of [0,n) given as (start, delta) lists:
// every piece of A is cut where its image crosses a piece boundary of B; delta_C = delta_A + delta_B.
// Pieces of A are independent: large inputs are cut into contiguous chunks, one host thread each.
static inline void compose_range(const std::vector<u64>& As, const std::vector<u64>& Ad, const std::vector<u64>& Bs,
                                 const std::vector<u64>& Bd, u64 n, u64 k0, u64 k1, std::vector<u64>& Cs,
                                 std::vector<u64>& Cd) {
    const u64 PA = As.size(), PB = Bs.size();
    Cs.reserve((k1 - k0) * 2 + 16); Cd.reserve((k1 - k0) * 2 + 16);
    for (u64 k = k0; k < k1; ++k) {
        const u64 s = As[k], e = (k + 1 < PA) ? As[k + 1] : n, a = Ad[k], len = e - s;
        u64 u = s + a;  // image of s
        if (u >= n) u -= n;
        u64 done = 0;
        while (done < len) {  // at most two linear ranges (the image may wrap around n)
            const bool wrapped = u + done >= n;
            const u64 pos = wrapped ? u + done - n : u + done;
            const u64 lin_end = wrapped ? (u + len - n) : std::min(u + len, n);
            u64 b = (u64)(std::upper_bound(Bs.begin(), Bs.end(), pos) - Bs.begin()) - 1;
            u64 cur = pos;
            while (cur < lin_end) {
                const u64 pend = std::min(lin_end, (b + 1 < PB) ? Bs[b + 1] : n);
                u64 d = a + Bd[b];
                if (d >= n) d -= n;
                Cs.push_back(s + done + (cur - pos));
                Cd.push_back(d);
                cur = pend; ++b;
            }
            done += lin_end - pos;
        }
    }
}

static inline void compose_translations(const std::vector<u64>& As, const std::vector<u64>& Ad,
                                        const std::vector<u64>& Bs, const std::vector<u64>& Bd, u64 n,
                                        std::vector<u64>& Cs, std::vector<u64>& Cd) {
    Cs.clear(); Cd.clear();
    const u64 PA = As.size();
    unsigned T = std::thread::hardware_concurrency();
    if (T > 32) T = 32;
    if (T < 2 || PA < (1u << 16)) { compose_range(As, Ad, Bs, Bd, n, 0, PA, Cs, Cd); return; }
    std::vector<std::vector<u64>> ps(T), pd(T);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t]() { compose_range(As, Ad, Bs, Bd, n, PA * t / T, PA * (t + 1) / T, ps[t], pd[t]); });
    for (auto& x : th) x.join();
    u64 total = 0;
    for (unsigned t = 0; t < T; ++t) total += ps[t].size();
    Cs.reserve(total); Cd.reserve(total);
    for (unsigned t = 0; t < T; ++t) {
        Cs.insert(Cs.end(), ps[t].begin(), ps[t].end());
        Cd.insert(Cd.end(), pd[t].begin(), pd[t].end());
        std::vector<u64>().swap(ps[t]); std::vector<u64>().swap(pd[t]);
    }
}

struct FlatHost {
    u64 n = 0, r = 0;
    u32 K = 16;          // runs per block
    u32 S = 0;           // distinct BWT symbols
    u64 nblk = 0;        // run blocks
    u32 lf_shift = 0;  u64 lf_nbkt = 0;
    u64 toe0 = 0;        // SA[n-1] = (samples_last[r-1]+1) % n, r_index.hpp:489
    std::vector<u64> F;            // [257]
    std::vector<uint16_t> sid;     // [256] dense symbol id or 0xFFFF
    std::vector<u64> start;        // [nblk*K + 1], padding = n
    std::vector<uint8_t> head;     // [nblk*K], padding = 0
    std::vector<u64> bstart;       // [nblk + 1], bstart[nblk] = n
    std::vector<u64> cum;          // [nblk*S*2] (count, last run id or ~0) — host-side staging only
    // Device form of the three arrays above, interleaved per block ("block record", one per K runs):
    //   [start x K : W-byte words][head x K : bytes][pad to W][count-before-block x S : W-byte words][pad to 32]
    // so that one rank query touches the directory sector (bdir) + ONE record (1-2 adjacent sectors)
    // instead of three separate arrays. W = 4 when w32 else 8.
    std::vector<uint8_t> blk;      // [nblk * blk_stride]
    u32 blk_stride = 0, off_head = 0, off_cum = 0;
    u32 rec_w = 8;       // bytes per position word INSIDE the block records: 4 (w32), 8, or 5 = 40-bit packed (n < 2^40)
    std::vector<u64> last;         // [nblk*S] id of the last run of each symbol before the block (locate toehold misses)
    std::vector<u32> bdir;         // [lf_nbkt + 1]
    std::vector<u64> samples_last; // [r]
    bool w32 = false;              // n < 2^32-1: Phi records and deltas are stored as 32-bit words
    PhiTable phi;                  // Phi^1..Phi^D refined
    // 64-bit index, D = 4, n < 2^40-1: the device gets 32-byte PACKED entries (4 x 40-bit deltas | 40-bit s1 or start |
    // 32-bit nxt | 24-bit cnt) instead of 8 x 8 bytes: one sector per lookup. With a table far beyond L2 (C5 at full
    // size: 493 MB) the window pass is DRAM-bound and half of its DRAM reads were the second sector of each entry.
    bool phi_packed = false;
    std::vector<uint8_t> phi_rec_p, phi_pent_p;
    JumpTable seed;                // Phi^SEG (seed.J = SEG; 0 = single-pass expansion only)
    u64 bytes() const {
        const u64 W = w32 ? 4 : 8;
        return F.size() * 8 + sid.size() * 2 + start.size() * W + blk.size() + bstart.size() * W + last.size() * W +
               bdir.size() * 4 + samples_last.size() * W + (phi_packed ? phi_rec_p.size() + phi_pent_p.size() : phi.bytes(w32)) + seed.bytes(w32);
    }
};

static inline u32 pick_shift(u64 n, u64 target_buckets) {
    if (target_buckets < 1) target_buckets = 1;
    u32 s = 0;
    while (((n - 1) >> s) + 1 > target_buckets) ++s;
    return s;
}

// Returns RIG_OK or RIG_ERR_INDEX / RIG_ERR_ARG. `max_bytes` (0 = unlimited) bounds the footprint.
static inline int flatten(const rig_logical_view& v, const rig_options& opt, FlatHost& f, u64 max_bytes = 0) {
    const bool timing = getenv("RIG_FLATTEN_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[flatten] %-28s %.3f s\n", what, std::chrono::duration<double>(t - t_last).count());
        t_last = t;
    };
    if (!v.F || !v.run_heads || !v.run_lens || !v.samples_last || !v.pred_pos || !v.pred_to_run) return RIG_ERR_ARG;
    if (v.n < 1 || v.r < 1 || v.r > v.n || v.r >= 0xFFFFFFF0ull) return RIG_ERR_INDEX;
    if (opt.runs_per_block != 0 && opt.runs_per_block != 4 && opt.runs_per_block != 8 && opt.runs_per_block != 16)
        return RIG_ERR_ARG;
    f.n = v.n; f.r = v.r;
    const u64 n = v.n, r = v.r;

    // symbols
    f.F.assign(v.F, v.F + 257);
    if (f.F[256] != n || f.F[0] != 0) return RIG_ERR_INDEX;
    for (int c = 0; c < 256; ++c) if (f.F[c] > f.F[c + 1]) return RIG_ERR_INDEX;
    f.sid.assign(256, 0xFFFF);
    {
        std::vector<char> seen(256, 0);
        for (u64 j = 0; j < r; ++j) seen[v.run_heads[j]] = 1;
        u32 S = 0;
        for (int c = 0; c < 256; ++c) if (seen[c]) f.sid[c] = (uint16_t)S++;
        f.S = S;
    }
    const u32 S = f.S;
    // Runs per block = lanes per rank query. Small groups put more patterns in a warp (the search
    // kernel is issue-bound: measured 0.50 / 0.30 / 0.19 ms for K = 16 / 8 / 4 on config C2), but the
    // per-block symbol directory costs S words per K runs (measured on the 1 GB pan-genome C4s, index
    // beyond L2: 6.1 / 8.5 / 14.5 ms for K = 4 / 8 / 16 — the kernel stays issue-bound even there).
    f.w32 = n < 0xFFFFFFFEull && !(opt.reserved[1] & 1);  // reserved[1] bit0: force 64-bit words (tests)
    const bool w32_pos = f.w32;
    u32 K = opt.runs_per_block;
    if (K == 0) {  // smallest K whose block records (K starts + K heads + S counts, 32-byte padded) stay within ~2 GB
        const u64 W = w32_pos ? 4 : (n < (1ull << 40) ? 5 : 8), budget = 2ull << 30;
        K = 16;
        for (u32 k : {4u, 8u}) {
            const u64 stride = (k * W + k + S * W + 3 + 31) / 32 * 32;
            if (((r + k - 1) / k) * stride <= budget) { K = k; break; }
        }
    }
    f.K = K;
    f.nblk = (r + K - 1) / K;
    const u64 nblk = f.nblk, rpad = nblk * K;
    if (max_bytes && nblk * S * 16 > max_bytes) return RIG_ERR_NOMEM;

    // run starts / heads / per-block directory
    f.start.assign(rpad + 1, n);
    f.head.assign(rpad, 0);
    f.bstart.assign(nblk + 1, n);
    f.cum.assign(nblk * (u64)S * 2, 0);
    std::vector<u64> cnt(S, 0), last(S, ~(u64)0);
    u64 pos = 0;
    for (u64 j = 0; j < r; ++j) {
        if (j % K == 0) {
            u64 b = j / K;
            f.bstart[b] = pos;
            for (u32 s = 0; s < S; ++s) { f.cum[(b * S + s) * 2] = cnt[s]; f.cum[(b * S + s) * 2 + 1] = last[s]; }
        }
        if (v.run_lens[j] == 0) return RIG_ERR_INDEX;
        if (j > 0 && v.run_heads[j] == v.run_heads[j - 1]) return RIG_ERR_INDEX;  // runs must be maximal
        f.start[j] = pos;
        f.head[j] = v.run_heads[j];
        u32 s = f.sid[v.run_heads[j]];
        cnt[s] += v.run_lens[j];
        last[s] = j;
        pos += v.run_lens[j];
        if (pos > n) return RIG_ERR_INDEX;
    }
    if (pos != n) return RIG_ERR_INDEX;
    for (int c = 0; c < 256; ++c) {  // F must agree with the runs
        u64 have = f.sid[c] == 0xFFFF ? 0 : cnt[f.sid[c]];
        if (f.F[c + 1] - f.F[c] != have) return RIG_ERR_INDEX;
    }

    lap("runs + directory counts");
    // interleaved block records
    {
        // Position words inside the records: 32-bit when w32; otherwise 40-bit packed while n < 2^40 (a K = 4 DNA
        // record is then 64 bytes = one DRAM burst instead of 96: the count kernel on a 10 GB text runs at the rate
        // DRAM delivers random sectors, 3 of its 4.4 sectors per rank query were the record), 64-bit beyond.
        // reserved[1] bit1 keeps 64-bit words (A/B switch).
        const u32 W = f.w32 ? 4 : ((n < (1ull << 40) && !(opt.reserved[1] & 2)) ? 5 : 8);
        f.rec_w = W;
        f.off_head = K * W;
        f.off_cum = W == 5 ? K * W + K : (K * W + K + W - 1) / W * W;
        f.blk_stride = (f.off_cum + S * W + (W == 5 ? 3 : 0) + 31) / 32 * 32;
        f.blk.assign(nblk * (u64)f.blk_stride, 0);
        f.last.assign(nblk * (u64)S, 0);
        auto put = [&](uint8_t* dst, u64 v) { memcpy(dst, &v, W); };  // little endian: the low W bytes
        for (u64 b = 0; b < nblk; ++b) {
            uint8_t* R = &f.blk[b * f.blk_stride];
            for (u32 t = 0; t < K; ++t) { put(R + t * W, f.start[b * K + t]); R[f.off_head + t] = f.head[b * K + t]; }
            for (u32 sy = 0; sy < S; ++sy) {
                put(R + f.off_cum + sy * W, f.cum[(b * S + sy) * 2]);
                f.last[b * S + sy] = f.cum[(b * S + sy) * 2 + 1];
            }
        }
    }

    // position -> block directory
    u32 fl = opt.lf_bucket_log2 ? opt.lf_bucket_log2 : 2;
    f.lf_shift = pick_shift(n, nblk << fl);
    f.lf_nbkt = ((n - 1) >> f.lf_shift) + 1;
    f.bdir.assign(f.lf_nbkt + 1, 0);
    {
        u64 b = 0;
        for (u64 q = 0; q < f.lf_nbkt; ++q) {
            u64 p = q << f.lf_shift;
            while (b + 1 < nblk && f.bstart[b + 1] <= p) ++b;
            f.bdir[q] = (u32)b;
        }
        f.bdir[f.lf_nbkt] = (u32)(nblk - 1);
    }

    lap("block records + bdir");
    // samples
    f.samples_last.assign(v.samples_last, v.samples_last + r);
    for (u64 j = 0; j < r; ++j) if (f.samples_last[j] >= n) return RIG_ERR_INDEX;
    f.toe0 = (f.samples_last[r - 1] + 1) % n;

    // Phi as a translation table. Samples p_0 < ... < p_{r-1} = n-1 (r_index.hpp:129); for
    // i in (p_k, p_{k+1}] the strict predecessor is p_k and Phi(i) = (samples_last[run_k - 1] + i - p_k) mod n
    // (r_index.hpp:205-219); for i in [0, p_0] the circular predecessor is p_{r-1} (:153-157, :210).
    // Pieces: [0, p_0] with the delta of sample r-1, then [p_k + 1, p_{k+1}] with the delta of sample k.
    PhiTable P1;
    {
        std::vector<u64> dl(r);
        for (u64 k = 0; k < r; ++k) {
            const u64 p = v.pred_pos[k], run = v.pred_to_run[k];
            if (p >= n || run >= r || (k > 0 && v.pred_pos[k - 1] >= p)) return RIG_ERR_INDEX;
            // run 0 is never reached by a valid query (Phi(SA[0]) is undefined, r_index.hpp:197,213)
            dl[k] = run > 0 ? (f.samples_last[run - 1] + n - p) % n : 0;
        }
        if (v.pred_pos[r - 1] != n - 1) return RIG_ERR_INDEX;  // last text position is always sampled
        P1.D = 1;
        P1.start.push_back(0); P1.delta.push_back(dl[r - 1]);
        for (u64 k = 0; k + 1 < r; ++k) { P1.start.push_back(v.pred_pos[k] + 1); P1.delta.push_back(dl[k]); }
    }
    // buckets per piece: 2 (measured on C2: 0.51 ms vs 0.57 ms with 4 — the smaller table stays in L2)
    const u32 fp = opt.phi_bucket_log2 ? opt.phi_bucket_log2 : 1;
    // D = occurrences produced per record lookup: requested, or the largest of {4,2,1} whose table fits the
    // byte budget (8 GB, or a quarter of the caller's limit). More occurrences per lookup win whether or not the
    // table stays in L2: measured on C5s (r = 3.4e5) D=4 at 93 MB 0.44 ms vs D=2 at 50 MB 0.94 ms; beyond L2 a
    // lookup is one random DRAM access (~55 G sectors/s) whatever it yields.
    u32 D = opt.reserved[0];
    if (D != 0 && D != 1 && D != 2 && D != 4 && D != 6 && D != 8) return RIG_ERR_ARG;
    if (D == 6 && !f.w32) return RIG_ERR_ARG;   // the 32-byte six-delta entry exists for 32-bit words only
    const bool auto_D = D == 0;
    if (D == 0) {
        u64 budget = 8ull << 30;
        if (max_bytes && max_bytes / 4 < budget) budget = max_bytes / 4;
        const u64 wb = f.w32 ? 4 : 8;
        const bool can_pack = !f.w32 && n < (1ull << 40) - 1 && !(opt.reserved[1] & 4);  // D = 4 entries packed into 32 bytes
        auto cost = [&](u32 d) { return (u64)d * r * (((d == 4 && can_pack) ? 32 : PhiTable::record_words(d) * wb) << fp); };
        // (D = 6 — six occurrences per 32-byte entry, 32-bit words — is available on request: a third of the lookups go
        // away, but measured on B200 it is no faster than D = 4 on C2 (0.461 vs 0.453 ms per step) and slower on C5s
        // (0.547 vs 0.455: the refined table grows by half and its lookups reach DRAM), so it is not the default.)
        D = cost(4) <= budget ? 4 : (cost(2) <= budget ? 2 : 1);
    }
    lap("Phi pieces");
    f.phi = P1;
    for (u32 j = 1; j < D; ++j) f.phi = extend_by_phi(f.phi, P1, n);
    lap("Phi^D composition");
    f.phi.build_directory(n, fp);
    if (D == 6 && (!f.phi.packable || f.phi.pieces() >= (1ull << 24))) {   // a crowded bucket or too many pieces: four per entry
        if (!auto_D) return RIG_ERR_ARG;
        D = 4;
        f.phi = P1;
        for (u32 j = 1; j < D; ++j) f.phi = extend_by_phi(f.phi, P1, n);
        f.phi.build_directory(n, fp);
    }
    f.phi_packed = false;
    if (!f.w32 && D == 4 && n < (1ull << 40) - 1 && !(opt.reserved[1] & 4)) {  // reserved[1] bit2: keep 8-byte words (A/B switch)
        bool fits = true;
        for (u64 q = 0; q < f.phi.nbkt && fits; ++q) fits = f.phi.rec[q * 8 + 6] < (1ull << 24);
        if (fits) {
            auto pack = [](const std::vector<u64>& src, std::vector<uint8_t>& dst) {
                const u64 cnt = src.size() / 8;
                dst.assign(cnt * 32, 0);
                for (u64 k = 0; k < cnt; ++k) {
                    const u64* w = &src[k * 8];
                    uint8_t* o = &dst[k * 32];
                    for (int t = 0; t < 5; ++t) { const u64 v = w[t] & 0xFFFFFFFFFFull; memcpy(o + 5 * t, &v, 5); }  // ~0 -> all ones
                    const uint32_t nxt = (uint32_t)w[5];
                    memcpy(o + 25, &nxt, 4);
                    const uint32_t c3 = (uint32_t)w[6];
                    memcpy(o + 29, &c3, 3);
                }
            };
            pack(f.phi.rec, f.phi_rec_p);
            pack(f.phi.pent, f.phi_pent_p);
            f.phi_packed = true;
        }
    }
    lap("Phi^D directory");
    if (f.phi.pieces() >= 0xFFFFFFF0ull) return RIG_ERR_INDEX;

    // Seed table Phi^SEG for the two-pass expansion (phi_kernels.cuh): requested (reserved[2] = 16..256,
    // 1 = off), or the largest of {128, 64, 32, 16} (measured on B200, round 2: 128 beats 64 by 4-8% per step on
    // C2 / C5s / C3s — half the dependent seed-table hops — and 256 loses: items of 256 slots balance badly) whose table (72 B per piece with 32-bit words: one
    // 64-byte bucket record per piece + an 8-byte piece entry) stays within 16 GB / a quarter of the
    // caller's byte limit. pieces(Phi^J) = sum over runs of min(J, run length), known before building.
    u32 SEG = opt.reserved[2];
    if (SEG != 0 && SEG != 1 && SEG != 16 && SEG != 32 && SEG != 64 && SEG != 128 && SEG != 256) return RIG_ERR_ARG;
    if (SEG == 0) {
        u64 budget = 16ull << 30;
        if (max_bytes && max_bytes / 4 < budget) budget = max_bytes / 4;
        const u64 per_piece = f.w32 ? 72 : 144;
        SEG = 1;
        for (u32 cand : {128u, 64u, 32u, 16u}) {
            u64 pieces = 0;
            for (u64 j = 0; j < r; ++j) pieces += std::min<u64>(cand, v.run_lens[j]);
            if (pieces * per_piece <= budget) { SEG = cand; break; }
        }
    }
    f.seed = JumpTable();
    if (SEG > 1) {
        std::vector<u64> s0 = P1.start, d0 = P1.delta, s1, d1;
        for (u32 J = 1; J < SEG; J *= 2) {
            compose_translations(s0, d0, s0, d0, n, s1, d1);
            s0.swap(s1); d0.swap(d1);
        }
        lap("seed composition (doubling)");
        f.seed.J = SEG;
        f.seed.start.swap(s0); f.seed.delta.swap(d0);
        if (f.seed.pieces() >= 0xFFFFFFF0ull) return RIG_ERR_INDEX;
        f.seed.build_directory(n, 0);
        lap("seed directory");
    }
    return RIG_OK;
}

}  // namespace rigf
