// logical_index.hpp — the layout-free content of an r-index and its
// construction from a text, plus the on-disk container written by ri-build.
//
// What the reference stores, and where (all under /root/reference/internal):
//   F[256]            r_index.hpp:72-82      #BWT symbols < c (terminator 0x01 included)
//   run heads/lengths rle_string.hpp:52-124  (there: run-boundary bitvectors + Huffman WT)
//   samples_last[j]   r_index.hpp:131-139    text position of the LAST symbol of run j, i.e. SA-1 (wrap to n-1)
//   pred / pred_to_run r_index.hpp:108-146   text positions of the FIRST symbol of every run, sorted, + run ids
// This header keeps the same information as plain arrays; the GPU library flattens them at load.
#pragma once
#include <exception>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>
#include <fstream>
#include <istream>
#include <ostream>
#include "sais.hpp"

namespace rib {

struct LogicalIndex {
    uint64_t n = 0;                    // BWT length = |T|+1            (r_index.hpp:88,454)
    uint64_t r = 0;                    // number of equal-letter runs   (r_index.hpp:92)
    uint64_t terminator_position = 0;  // x with BWT[x]=0x01            (r_index.hpp:84-86)
    uint64_t F[257] = {0};             // F[256] = n (the reference reads F[c+1] for c=255)
    std::vector<uint8_t> run_heads;    // [r]
    std::vector<uint64_t> run_lens;    // [r]
    std::vector<uint64_t> samples_last;  // [r] in BWT-run order
    std::vector<uint64_t> pred_pos;      // [r] ascending text positions of run-first samples
    std::vector<uint64_t> pred_to_run;   // [r] run id of pred_pos[k]
};

static const uint8_t kTerminator = 1;  // r_index.hpp:646

inline bool contains_reserved_chars(const uint8_t* t, uint64_t len) {  // r_index.hpp:636-644
    for (uint64_t i = 0; i < len; ++i)
        if (t[i] == 0 || t[i] == 1) return true;
    return false;
}

// Scan SA once (the reference streams it from SDSL's cache file, r_index.hpp:575-623) and emit
// runs + both sample sets. I is the SA index type.
template <class I>
inline void runs_and_samples_from_sa(const uint8_t* text, uint64_t n, const I* SA, LogicalIndex& L) {
    L.n = n;
    L.run_heads.clear(); L.run_lens.clear(); L.samples_last.clear();
    std::vector<std::pair<uint64_t, uint64_t>> first;  // (text position, run id)
    uint64_t hist[256] = {0};
    auto smp = [&](uint64_t x) -> uint64_t { uint64_t v = (uint64_t)SA[x]; return v > 0 ? v - 1 : n - 1; };
    auto bwt = [&](uint64_t x) -> uint8_t { uint64_t v = (uint64_t)SA[x]; return v > 0 ? text[v - 1] : kTerminator; };
    uint8_t cur = bwt(0);
    uint64_t run_start = 0;
    first.push_back({smp(0), 0});
    for (uint64_t x = 1; x <= n; ++x) {
        uint8_t c = (x < n) ? bwt(x) : 0;
        if (x == n || c != cur) {
            L.run_heads.push_back(cur);
            L.run_lens.push_back(x - run_start);
            L.samples_last.push_back(smp(x - 1));
            hist[cur] += x - run_start;
            if (cur == kTerminator) L.terminator_position = run_start;
            if (x < n) { first.push_back({smp(x), (uint64_t)L.run_heads.size()}); cur = c; run_start = x; }
        }
    }
    L.r = L.run_heads.size();
    uint64_t acc = 0;
    for (int c = 0; c < 256; ++c) { L.F[c] = acc; acc += hist[c]; }
    L.F[256] = n;
    std::sort(first.begin(), first.end());  // r_index.hpp:108
    L.pred_pos.resize(L.r); L.pred_to_run.resize(L.r);
    for (uint64_t k = 0; k < L.r; ++k) { L.pred_pos[k] = first[k].first; L.pred_to_run[k] = first[k].second; }
}

// Build from a text (no 0x00/0x01 bytes). Throws std::invalid_argument on reserved bytes
// (the CLI maps that to the reference's message + exit(1), r_index.hpp:46-51).
inline LogicalIndex build_logical_index(const uint8_t* text, uint64_t len) {
    if (contains_reserved_chars(text, len)) throw std::invalid_argument("reserved");
    LogicalIndex L;
    uint64_t n = len + 1;
    std::vector<uint8_t> buf(n);
    if (len) std::memcpy(buf.data(), text, len);
    buf[len] = 0;
    if (n < (uint64_t(1) << 31) - 2) {
        std::vector<int32_t> SA(n);
        suffix_array_with_sentinel<int32_t>(buf.data(), (int32_t)n, SA.data());
        runs_and_samples_from_sa<int32_t>(buf.data(), n, SA.data(), L);
    } else {
        std::vector<int64_t> SA(n);
        suffix_array_with_sentinel<int64_t>(buf.data(), (int64_t)n, SA.data());
        runs_and_samples_from_sa<int64_t>(buf.data(), n, SA.data(), L);
    }
    return L;
}

// ---- container (.ri written by THIS repo's ri-build) --------------------------------------
// The reference's .ri embeds SDSL-private blobs (out of scope); ours is a
// versioned flat container. Like the reference, the CLI writes one `fast` flag byte first
// (ri-build.cpp:133) and the readers skip it (ri-count.cpp:155-158).
static const char kMagic[8] = {'R', 'I', 'B', '2', '0', '0', 'v', '1'};

inline uint64_t serialize(const LogicalIndex& L, std::ostream& out) {
    uint64_t w = 0;
    auto put = [&](const void* p, size_t b) { out.write((const char*)p, (std::streamsize)b); w += b; };
    put(kMagic, 8);
    put(&L.n, 8); put(&L.r, 8); put(&L.terminator_position, 8);
    put(L.F, 257 * 8);
    put(L.run_heads.data(), L.r);
    put(L.run_lens.data(), L.r * 8);
    put(L.samples_last.data(), L.r * 8);
    put(L.pred_pos.data(), L.r * 8);
    put(L.pred_to_run.data(), L.r * 8);
    return w;
}

inline bool load(LogicalIndex& L, std::istream& in) {
    char mg[8];
    in.read(mg, 8);
    if (!in || std::memcmp(mg, kMagic, 8) != 0) return false;
    in.read((char*)&L.n, 8); in.read((char*)&L.r, 8); in.read((char*)&L.terminator_position, 8);
    in.read((char*)L.F, 257 * 8);
    if (!in) return false;
    // sanity before any allocation: a truncated or corrupt file must be refused, not turned into a bad_alloc
    if (L.n < 1 || L.r < 1 || L.r > L.n || L.terminator_position >= L.n || L.F[0] != 0 || L.F[256] != L.n) return false;
    {
        const std::streampos here = in.tellg();
        if (here != std::streampos(-1)) {   // seekable: the five arrays (33 bytes per run) must all be there
            in.seekg(0, std::ios::end);
            const std::streampos end = in.tellg();
            in.seekg(here);
            if (end == std::streampos(-1) || (uint64_t)(end - here) / 33 < L.r) return false;
        }
    }
    try {
        L.run_heads.resize(L.r); L.run_lens.resize(L.r); L.samples_last.resize(L.r);
        L.pred_pos.resize(L.r); L.pred_to_run.resize(L.r);
    } catch (const std::exception&) {
        return false;
    }
    in.read((char*)L.run_heads.data(), (std::streamsize)L.r);
    in.read((char*)L.run_lens.data(), (std::streamsize)(L.r * 8));
    in.read((char*)L.samples_last.data(), (std::streamsize)(L.r * 8));
    in.read((char*)L.pred_pos.data(), (std::streamsize)(L.r * 8));
    in.read((char*)L.pred_to_run.data(), (std::streamsize)(L.r * 8));
    return (bool)in;
}

}  // namespace rib
