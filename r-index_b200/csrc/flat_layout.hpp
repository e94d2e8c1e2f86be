// flat_layout.hpp — "flatten once at load": logical r-index arrays -> the word arrays the
// kernels read. Plain C++ (no CUDA), so the layout can also be checked on a CPU-only box by
// tests/support/flat_check.cpp (a test double that walks these same arrays with scalar code).
//
// What the reference keeps as SDSL objects, and what replaces each here:
//
//   reference member (internal/…)                      flat arrays (all O(r) words)
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
//   r_index::pred (EF) + pred_to_run + samples_last      phi_ent[]: (text position, delta) pairs sorted
//     r_index.hpp:663-665                                  by position, delta = samples_last[run-1] - pos
//                                                         (mod n), so Phi(i) = (i + delta) mod n;
//                                                       phi_dir[]: direct-addressed position>>s -> rank
//   r_index::samples_last                               samples_last[] : u64, run order (toeholds,
//                                                         chain splitting at run boundaries)
//   r_index::F                                          F[257], sid[256] (symbol -> dense id)
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>
#include "../../include/rindex_gpu.h"

namespace rigf {

typedef uint64_t u64;
typedef uint32_t u32;

struct FlatHost {
    u64 n = 0, r = 0;
    u32 K = 16;          // runs per block
    u32 S = 0;           // distinct BWT symbols
    u64 nblk = 0;        // run blocks
    u32 lf_shift = 0;  u64 lf_nbkt = 0;
    u32 phi_shift = 0; u64 phi_nbkt = 0;
    u64 toe0 = 0;        // SA[n-1] = (samples_last[r-1]+1) % n, r_index.hpp:489
    std::vector<u64> F;            // [257]
    std::vector<uint16_t> sid;     // [256] dense symbol id or 0xFFFF
    std::vector<u64> start;        // [nblk*K + 1], padding = n
    std::vector<uint8_t> head;     // [nblk*K], padding = 0
    std::vector<u64> bstart;       // [nblk + 1], bstart[nblk] = n
    std::vector<u64> cum;          // [nblk*S*2] (count, last run id or ~0)
    std::vector<u32> bdir;         // [lf_nbkt + 1]
    std::vector<u64> samples_last; // [r]
    std::vector<u64> phi_ent;      // [2r] (pos, delta)
    std::vector<u32> phi_dir;      // [phi_nbkt + 1]
    u64 bytes() const {
        return F.size() * 8 + sid.size() * 2 + start.size() * 8 + head.size() + bstart.size() * 8 + cum.size() * 8 +
               bdir.size() * 4 + samples_last.size() * 8 + phi_ent.size() * 8 + phi_dir.size() * 4;
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
    if (!v.F || !v.run_heads || !v.run_lens || !v.samples_last || !v.pred_pos || !v.pred_to_run) return RIG_ERR_ARG;
    if (v.n < 1 || v.r < 1 || v.r > v.n || v.r >= 0xFFFFFFF0ull) return RIG_ERR_INDEX;
    u32 K = opt.runs_per_block ? opt.runs_per_block : 16;
    if (K != 4 && K != 8 && K != 16) return RIG_ERR_ARG;
    f.n = v.n; f.r = v.r; f.K = K;
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

    // samples + Phi
    f.samples_last.assign(v.samples_last, v.samples_last + r);
    for (u64 j = 0; j < r; ++j) if (f.samples_last[j] >= n) return RIG_ERR_INDEX;
    f.toe0 = (f.samples_last[r - 1] + 1) % n;
    f.phi_ent.assign(2 * r, 0);
    for (u64 k = 0; k < r; ++k) {
        u64 p = v.pred_pos[k], run = v.pred_to_run[k];
        if (p >= n || run >= r || (k > 0 && v.pred_pos[k - 1] >= p)) return RIG_ERR_INDEX;
        f.phi_ent[2 * k] = p;
        // Phi(i) = (samples_last[run-1] + (i - p)) mod n  (r_index.hpp:205-219; wrap case :210 is the
        // same formula mod n because the circular predecessor is then n-1). run 0 is never reached
        // (Phi(SA[0]) is undefined, r_index.hpp:197,213).
        f.phi_ent[2 * k + 1] = run > 0 ? (f.samples_last[run - 1] + n - p) % n : 0;
    }
    if (v.pred_pos[r - 1] != n - 1) return RIG_ERR_INDEX;  // last text position is always sampled, r_index.hpp:129
    u32 fp = opt.phi_bucket_log2 ? opt.phi_bucket_log2 : 2;
    f.phi_shift = pick_shift(n, r << fp);
    f.phi_nbkt = ((n - 1) >> f.phi_shift) + 1;
    f.phi_dir.assign(f.phi_nbkt + 1, 0);
    {
        u64 k = 0;
        for (u64 q = 0; q < f.phi_nbkt; ++q) {
            u64 p = q << f.phi_shift;
            while (k < r && v.pred_pos[k] < p) ++k;
            f.phi_dir[q] = (u32)k;
        }
        f.phi_dir[f.phi_nbkt] = (u32)r;
    }
    return RIG_OK;
}

}  // namespace rigf
