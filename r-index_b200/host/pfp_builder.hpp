// pfp_builder.hpp — scalable index construction by PREFIX-FREE PARSING
// (Boucher, Gagie, Kuhnle, Langmead, Manzini, Mun: "Prefix-free parsing for building big BWTs", 2019).
//
// Replaces, for large repetitive inputs, the suffix-array route of the reference's constructor
// (internal/r_index.hpp:42-150 with sufsort :553-634: construct_sa over the whole text, then one scan of SA
// that emits BWT runs and both sample sets). The in-memory SA-IS builder (logical_index.hpp) needs 5-9
// bytes per text symbol; this one needs memory proportional to the DICTIONARY and the PARSE of the text:
//
//   S = T·0x00 is cut into overlapping phrases at every window of w bytes whose Karp-Rabin hash is 0 mod p
//   (each phrase starts with the trigger window that ended the previous one). The distinct phrases form
//   the dictionary D (sorted, ranks = new alphabet), the sequence of ranks the parse P.
//   Every suffix of S starts inside exactly one phrase occurrence at an offset that leaves more than w
//   bytes of the phrase; those phrase suffixes form a prefix-free set, so two text suffixes compare
//   (1) by their phrase suffixes, and, when these are the same string, (2) by the parse suffixes that
//   follow — i.e. by the suffix array of P.
//   So: suffix array of P (|P| ~ n/p integers), suffix array of D (a few bytes per dictionary byte), one
//   sweep over D's suffixes in order. A group of equal phrase suffixes that are all preceded by the same
//   byte is ONE block of the BWT (count = number of occurrences of its phrases): O(1) work however many
//   text positions it covers; only groups with different preceding bytes (the neighbourhood of BWT run
//   boundaries) are merged occurrence by occurrence.
//
// The sweep feeds the same run/sample logic as the SA scan (terminator 0x01 in the row with SA = 0,
// samples = SA-1 with wrap to n-1, run-first samples sorted by text position: r_index.hpp:587-623,
// :108,:141-146), so both builders produce identical LogicalIndex arrays (tests/test_host.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <chrono>
#include <memory>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>
#include "logical_index.hpp"

namespace rib {
namespace pfp {

struct Params {
    uint32_t w = 10;   // window (overlap) length
    uint32_t p = 100;  // a window triggers when hash mod p == 0: phrases average ~p bytes
};

struct Stats {
    uint64_t phrases = 0, dict_bytes = 0, parse_len = 0, groups = 0, uniform_rows = 0, merged_rows = 0;
};

// Consumer of BWT rows in suffix-array order: (symbol, number of consecutive rows, SA of the first and last row).
class RunBuilder {
public:
    RunBuilder(LogicalIndex& L_, uint64_t n_) : L(L_), n(n_) {
        L.run_heads.clear(); L.run_lens.clear(); L.samples_last.clear();
        std::memset(hist, 0, sizeof(hist));
    }
    inline void emit(uint8_t c, uint64_t count, uint64_t sa_first, uint64_t sa_last) {
        if (!open || c != cur) {
            if (open) close();
            open = true; cur = c; run_len = 0;
            first.push_back({smp(sa_first), (uint64_t)L.run_heads.size()});
            if (c == kTerminator) L.terminator_position = rows;
        }
        run_len += count; rows += count; last_sa = sa_last;
    }
    uint64_t rows_emitted() const { return rows; }
    void finish() {
        if (open) close();
        L.n = n;
        L.r = L.run_heads.size();
        uint64_t acc = 0;
        for (int c = 0; c < 256; ++c) { L.F[c] = acc; acc += hist[c]; }
        L.F[256] = n;
        std::sort(first.begin(), first.end());  // r_index.hpp:108
        L.pred_pos.resize(L.r); L.pred_to_run.resize(L.r);
        for (uint64_t k = 0; k < L.r; ++k) { L.pred_pos[k] = first[k].first; L.pred_to_run[k] = first[k].second; }
    }

private:
    inline uint64_t smp(uint64_t sa) const { return sa > 0 ? sa - 1 : n - 1; }  // r_index.hpp:599,604,614,619
    void close() {
        L.run_heads.push_back(cur); L.run_lens.push_back(run_len); L.samples_last.push_back(smp(last_sa));
        hist[cur] += run_len;
    }
    LogicalIndex& L;
    uint64_t n;
    std::vector<std::pair<uint64_t, uint64_t>> first;  // (text position, run id)
    uint64_t hist[256];
    bool open = false;
    uint8_t cur = 0;
    uint64_t run_len = 0, rows = 0, last_sa = 0;
};

// dictionary symbols: 0 = end of the concatenation (SA-IS sentinel), 1 = phrase separator, 2 = the 0x00 that ends S,
// text byte b (>= 2) -> b + 1
static inline uint16_t sym_of(uint8_t b) { return b == 0 ? (uint16_t)2 : (uint16_t)(b + 1); }
static inline uint8_t byte_of(uint16_t s) { return s == 2 ? (uint8_t)0 : (uint8_t)(s - 1); }

inline LogicalIndex build(const uint8_t* text, uint64_t len, const Params& prm = Params(), Stats* stats = nullptr) {
    if (contains_reserved_chars(text, len)) throw std::invalid_argument("reserved");
    const uint64_t n = len + 1;  // S = T·0x00
    const uint32_t w = prm.w < 1 ? 1 : prm.w, p = prm.p < 1 ? 1 : prm.p;
    auto S = [&](uint64_t i) -> uint8_t { return i < len ? text[i] : (uint8_t)0; };
    const bool timing = getenv("RIB_PFP_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[pfp] %-34s %.2f s\n", what, std::chrono::duration<double>(t - t_last).count());
        t_last = t;
    };

    // ---- 1. parse -------------------------------------------------------------------------------------------
    // (a) trigger positions: a window S[i-w+1..i] triggers when a mix of its rolling polynomial hash (mod 2^64)
    //     falls below 2^32/p. A pure function of the window: text chunks are scanned by independent host threads.
    // (b) phrase k = S[E[k-1]-w+1 .. E[k]] (the first starts at 0; the first cut needs i >= w so that the phrase is
    //     longer than the window); (c) per-phrase content hashes, again in parallel; (d) dictionary, serial:
    //     |P| ~ n/p hash lookups, each verified byte for byte. Phrases are kept as (offset, length) into the text;
    //     only the last one, which ends with S's 0x00, is materialised.
    struct Phrase { uint64_t off; uint32_t len; };    // bytes S[off .. off+len)
    std::vector<Phrase> phrases;                       // by id (first-seen order)
    std::vector<uint32_t> parse;                       // phrase ids
    std::vector<uint64_t> tpos;                        // start of every phrase occurrence in S
    std::unordered_multimap<uint64_t, uint32_t> seen;  // content hash -> phrase id (content verified)
    {
        unsigned T = std::thread::hardware_concurrency();
        if (T > 32) T = 32;
        if (T < 1 || n < ((uint64_t)1 << 22)) T = 1;
        if (const char* e = getenv("RIB_PFP_THREADS")) T = std::max(1, std::min(64, atoi(e)));  // tests: force the threaded paths
        const uint64_t B = 0x100000001b3ull;
        uint64_t Bw = 1;
        for (uint32_t t = 0; t < w; ++t) Bw *= B;
        const uint64_t thr = ((uint64_t)1 << 32) / p;
        std::vector<std::vector<uint64_t>> trig(T);
        auto scan = [&](unsigned t) {
            const uint64_t c0 = n * t / T, c1 = n * (t + 1) / T, i0 = c0 >= w ? c0 - w : 0;
            uint64_t hw = 0;
            std::vector<uint64_t>& out = trig[t];
            for (uint64_t i = i0; i < c1; ++i) {
                hw = hw * B + (uint64_t)S(i) + 1;
                if (i - i0 >= w) hw -= Bw * ((uint64_t)S(i - w) + 1);
                if (i < c0 || i < w) continue;  // i >= w: the window is complete and the first phrase is longer than it
                const uint64_t mix = (hw ^ (hw >> 29)) * 0xBF58476D1CE4E5B9ull;
                if ((mix >> 32) < thr) out.push_back(i);
            }
        };
        if (T == 1) scan(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t) th.emplace_back(scan, t);
            for (auto& x : th) x.join();
        }
        lap("parse: trigger scan");
        std::vector<uint64_t> E;  // inclusive end of every phrase
        {
            uint64_t total = 1;
            for (auto& v : trig) total += v.size();
            E.reserve(total);
            for (auto& v : trig) { E.insert(E.end(), v.begin(), v.end()); std::vector<uint64_t>().swap(v); }
            if (E.empty() || E.back() != n - 1) E.push_back(n - 1);
        }
        const uint64_t NPh = E.size();
        if (NPh >= 0x7fffff00ull) throw std::length_error("parse longer than 2^31");
        tpos.resize(NPh);
        std::vector<uint64_t> phash(NPh);
        auto hash_range = [&](unsigned t) {
            for (uint64_t k = NPh * t / T; k < NPh * (t + 1) / T; ++k) {
                const uint64_t b0 = k == 0 ? 0 : E[k - 1] + 1 - w, e0 = E[k];
                if (e0 - b0 + 1 > 0xffffffffull) continue;  // reported by the serial pass
                uint64_t h = 1469598103934665603ull;
                for (uint64_t i = b0; i <= e0; ++i) h = (h ^ S(i)) * 1099511628211ull;
                tpos[k] = b0; phash[k] = h;
            }
        };
        if (T == 1) hash_range(0);
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t) th.emplace_back(hash_range, t);
            for (auto& x : th) x.join();
        }
        lap("parse: phrase hashes");
        parse.resize(NPh);
        for (uint64_t k = 0; k < NPh; ++k) {
            const uint64_t b0 = k == 0 ? 0 : E[k - 1] + 1 - w, L64 = E[k] - b0 + 1;
            if (L64 > 0xffffffffull) throw std::length_error("phrase longer than 2^32");
            const uint32_t L = (uint32_t)L64;
            uint32_t id = ~0u;
            if (k + 1 < NPh) {  // the last phrase holds the unique 0x00: always new
                auto range = seen.equal_range(phash[k]);
                for (auto it = range.first; it != range.second; ++it) {
                    const Phrase& q = phrases[it->second];
                    if (q.len == L && std::memcmp(text + q.off, text + b0, L) == 0) { id = it->second; break; }
                }
            }
            if (id == ~0u) {
                id = (uint32_t)phrases.size();
                if (phrases.size() >= 0x7fffff00ull) throw std::length_error("too many distinct phrases");
                phrases.push_back({b0, L});
                if (k + 1 < NPh) seen.emplace(phash[k], id);
            }
            parse[k] = id;
        }
    }
    seen.clear();
    lap("parse: dictionary (serial)");
    const uint64_t ND = phrases.size(), NP = parse.size();
    if (NP >= 0x7fffff00ull) throw std::length_error("parse longer than 2^31");

    // ---- 2. sort the dictionary, rename the parse --------------------------------------------------------------
    std::vector<uint32_t> order(ND);
    for (uint32_t i = 0; i < ND; ++i) order[i] = i;
    // phrase bytes: text[off .. off+len), with the virtual 0x00 of S as the last byte of the last phrase only
    const uint32_t last_id = parse[NP - 1];
    auto pbyte = [&](const Phrase& q, uint32_t t) -> uint8_t { return q.off + t < len ? text[q.off + t] : (uint8_t)0; };
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const Phrase &x = phrases[a], &y = phrases[b];
        uint32_t m = std::min(x.len, y.len);
        if (a == last_id || b == last_id) {  // compare byte-wise through the virtual terminator
            for (uint32_t t = 0; t < m; ++t) {
                const uint8_t cx = pbyte(x, t), cy = pbyte(y, t);
                if (cx != cy) return cx < cy;
            }
            return x.len < y.len;
        }
        const int c = std::memcmp(text + x.off, text + y.off, m);
        return c != 0 ? c < 0 : x.len < y.len;
    });
    std::vector<uint32_t> rank_of(ND);
    for (uint32_t i = 0; i < ND; ++i) rank_of[order[i]] = i;
    for (auto& x : parse) x = rank_of[x];
    const uint32_t last_d = parse[NP - 1];  // the phrase that ends with S's 0x00: unique, occurs once

    // dictionary in rank order as 16-bit symbols with separators
    std::vector<uint64_t> dstart(ND + 1);
    std::vector<uint32_t> dlen(ND);
    uint64_t M = 0;
    for (uint32_t d = 0; d < ND; ++d) { dstart[d] = M; dlen[d] = phrases[order[d]].len; M += (uint64_t)dlen[d] + 1; }
    dstart[ND] = M;
    std::vector<uint16_t> C(M + 1);
    for (uint32_t d = 0; d < ND; ++d) {
        const Phrase& q = phrases[order[d]];
        uint16_t* dst = &C[dstart[d]];
        for (uint32_t t = 0; t < q.len; ++t) dst[t] = sym_of(pbyte(q, t));
        dst[q.len] = 1;
    }
    C[M] = 0;
    phrases.clear(); phrases.shrink_to_fit(); order.clear(); order.shrink_to_fit();
    rank_of.clear(); rank_of.shrink_to_fit();

    lap("dictionary sort + rename");
    // ---- 3. suffix array of the parse; occurrences of every phrase ordered by the parse suffix that follows --------
    std::vector<uint32_t> occ_begin(ND + 1, 0), occ_rank(NP), occ_k(NP);
    {
        std::vector<uint32_t> P1(NP + 1);
        for (uint64_t k = 0; k < NP; ++k) P1[k] = parse[k] + 1;
        P1[NP] = 0;
        std::vector<int32_t> SAP(NP + 1);
        sais_detail::sais_rec<uint32_t, int32_t>(P1.data(), SAP.data(), (int32_t)(NP + 1), (int32_t)ND);
        for (uint64_t k = 0; k < NP; ++k) occ_begin[parse[k] + 1]++;
        for (uint32_t d = 0; d < ND; ++d) occ_begin[d + 1] += occ_begin[d];
        std::vector<uint32_t> fill(occ_begin.begin(), occ_begin.end() - 1);
        for (uint64_t i = 0; i <= NP; ++i) {  // i = rank of the parse suffix starting at SAP[i]
            const int32_t j = SAP[i];
            if (j < 1) continue;
            const uint32_t k = (uint32_t)j - 1, d = parse[k];
            occ_rank[fill[d]] = (uint32_t)i; occ_k[fill[d]] = k; ++fill[d];
        }
    }

    lap("suffix array of the parse + lists");
    // ---- 4. suffix array of the dictionary -----------------------------------------------------------------------
    if (M + 1 >= 0x7fffff00ull) throw std::length_error("dictionary longer than 2^31");
    std::vector<int32_t> SAD(M + 1);
    sais_detail::sais_rec<uint16_t, int32_t>(C.data(), SAD.data(), (int32_t)(M + 1), (int32_t)257);

    lap("suffix array of the dictionary");
    // ---- 5. sweep ------------------------------------------------------------------------------------------------
    // The dictionary's suffix array is cut at group boundaries into one range per host thread; every thread emits
    // the BWT runs of its range into a private list, and the lists are fed in order to the run/sample builder.
    LogicalIndex L;
    RunBuilder rb(L, n);
    Stats st;
    st.phrases = ND; st.dict_bytes = M; st.parse_len = NP;
    struct Member { uint32_t d, o; };
    struct RunList {  // maximal runs of one range: symbol, length, SA of the first and last row
        std::vector<uint8_t> c; std::vector<uint64_t> cnt, sf, sl;
        inline void emit(uint8_t ch, uint64_t count, uint64_t sa_first, uint64_t sa_last) {
            if (!c.empty() && c.back() == ch) { cnt.back() += count; sl.back() = sa_last; }
            else { c.push_back(ch); cnt.push_back(count); sf.push_back(sa_first); sl.push_back(sa_last); }
        }
    };
    auto pred_byte_of_occurrence = [&](uint32_t k) -> uint8_t {  // the byte of S before phrase occurrence k
        if (k == 0) return kTerminator;                           // the row with SA = 0 (r_index.hpp:587-590)
        const uint32_t pd = parse[k - 1];
        return byte_of(C[dstart[pd] + dlen[pd] - w - 1]);
    };
    // a dictionary suffix that stands for text suffixes: not a separator, and longer than the overlap (or in the last phrase)
    auto decode = [&](uint64_t idx, uint64_t& pos, uint32_t& dd, uint32_t& oo, uint64_t& rem) -> bool {
        pos = (uint64_t)SAD[idx];
        if (C[pos] == 1) return false;
        dd = (uint32_t)(std::upper_bound(dstart.begin(), dstart.begin() + ND, pos) - dstart.begin()) - 1;
        oo = (uint32_t)(pos - dstart[dd]);
        rem = dlen[dd] - oo;
        return rem > w || dd == last_d;
    };
    auto same_string = [&](uint64_t pa, uint64_t ra, uint64_t pb, uint64_t rbm) {
        return ra == rbm && std::memcmp(&C[pa], &C[pb], ra * sizeof(uint16_t)) == 0;
    };
    auto sweep_range = [&](uint64_t i0, uint64_t i1, RunList& sink, Stats& lst) {
        std::vector<Member> grp;
        auto flush_group = [&]() {
            if (grp.empty()) return;
            ++lst.groups;
            bool uniform = true;
            uint8_t c0 = 0;
            bool have = false;
            for (const Member& m : grp) {
                if (m.o == 0) { uniform = false; break; }
                const uint8_t c = byte_of(C[dstart[m.d] + m.o - 1]);
                if (!have) { c0 = c; have = true; } else if (c != c0) { uniform = false; break; }
            }
            if (uniform) {
                uint64_t total = 0, sa_first = 0, sa_last = 0;
                uint32_t rmin = ~0u, rmax = 0;
                bool any = false;
                for (const Member& m : grp) {
                    const uint32_t a = occ_begin[m.d], b = occ_begin[m.d + 1];
                    if (a == b) continue;
                    total += b - a;
                    if (!any || occ_rank[a] < rmin) { rmin = occ_rank[a]; sa_first = tpos[occ_k[a]] + m.o; }
                    if (!any || occ_rank[b - 1] > rmax) { rmax = occ_rank[b - 1]; sa_last = tpos[occ_k[b - 1]] + m.o; }
                    any = true;
                }
                if (total) { sink.emit(c0, total, sa_first, sa_last); lst.uniform_rows += total; }
            } else if (grp.size() == 1) {
                const Member m = grp[0];
                for (uint32_t x = occ_begin[m.d]; x < occ_begin[m.d + 1]; ++x) {
                    const uint32_t k = occ_k[x];
                    const uint64_t sa = tpos[k] + m.o;
                    sink.emit(m.o > 0 ? byte_of(C[dstart[m.d] + m.o - 1]) : pred_byte_of_occurrence(k), 1, sa, sa);
                    ++lst.merged_rows;
                }
            } else {  // merge the members' occurrence lists by the rank of the following parse suffix
                typedef std::pair<uint32_t, uint32_t> QE;  // (rank, member index)
                std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
                std::vector<uint32_t> cur(grp.size());
                for (uint32_t g = 0; g < grp.size(); ++g) {
                    cur[g] = occ_begin[grp[g].d];
                    if (cur[g] < occ_begin[grp[g].d + 1]) pq.push({occ_rank[cur[g]], g});
                }
                // Pop the list with the smallest next rank and drain it up to the next list's head: the occurrences of one
                // phrase suffix often come in long stretches of consecutive ranks, and when the suffix does not start
                // its phrase (o > 0) they all carry the same BWT symbol — a whole stretch is one emit().
                while (!pq.empty()) {
                    const uint32_t g = pq.top().second;
                    pq.pop();
                    const uint32_t limit = pq.empty() ? ~0u : pq.top().first;  // ranks are distinct across lists
                    const Member m = grp[g];
                    const uint32_t end = occ_begin[m.d + 1];
                    uint32_t x = cur[g];
                    if (m.o > 0) {
                        // gallop to the first occurrence whose rank exceeds the limit: everything below lo2 is known to be
                        // under it, hi2 is `end` or the first probe at or above it
                        uint32_t lo2 = x + 1, hi2 = x + 1, step = 1;
                        while (hi2 < end && occ_rank[hi2] < limit) {
                            lo2 = hi2 + 1;
                            hi2 = (end - hi2 > step) ? hi2 + step : end;
                            step <<= 1;
                        }
                        const uint32_t stop = (uint32_t)(std::lower_bound(occ_rank.begin() + lo2, occ_rank.begin() + hi2, limit) - occ_rank.begin());
                        sink.emit(byte_of(C[dstart[m.d] + m.o - 1]), stop - x, tpos[occ_k[x]] + m.o, tpos[occ_k[stop - 1]] + m.o);
                        lst.merged_rows += stop - x;
                        x = stop;
                    } else {
                        do {
                            const uint32_t k = occ_k[x];
                            sink.emit(pred_byte_of_occurrence(k), 1, tpos[k], tpos[k]);
                            ++lst.merged_rows;
                            ++x;
                        } while (x < end && occ_rank[x] < limit);
                    }
                    cur[g] = x;
                    if (x < end) pq.push({occ_rank[x], g});
                }
            }
            grp.clear();
        };
        uint64_t rep_pos = 0, rep_rem = 0;  // representative suffix of the current group
        for (uint64_t idx = i0; idx < i1; ++idx) {
            uint64_t pos, rem; uint32_t dd, oo;
            if (!decode(idx, pos, dd, oo, rem)) continue;
            if (!(!grp.empty() && same_string(pos, rem, rep_pos, rep_rem))) { flush_group(); rep_pos = pos; rep_rem = rem; }
            grp.push_back({dd, oo});
        }
        flush_group();
    };
    unsigned T = std::thread::hardware_concurrency();
    if (T > 32) T = 32;
    if (T < 1 || M < ((uint64_t)1 << 20)) T = 1;
    if (const char* e = getenv("RIB_PFP_THREADS")) T = std::max(1, std::min(64, atoi(e)));
    // range t starts at the first group boundary at or after 1 + M*t/T (SAD[0] is the end-of-concatenation sentinel)
    std::vector<uint64_t> cut(T + 1, M + 1);
    cut[0] = 1;
    for (unsigned t = 1; t < T; ++t) {
        uint64_t idx = 1 + M * t / T;
        uint64_t ppos = 0, prem = 0; bool have_prev = false;
        for (uint64_t b = idx; b-- > 1;) {  // the valid suffix before idx, if any
            uint64_t pos, rem; uint32_t dd, oo;
            if (decode(b, pos, dd, oo, rem)) { ppos = pos; prem = rem; have_prev = true; break; }
        }
        for (; idx <= M; ++idx) {
            uint64_t pos, rem; uint32_t dd, oo;
            if (!decode(idx, pos, dd, oo, rem)) continue;
            if (!have_prev || !same_string(pos, rem, ppos, prem)) break;  // a group starts here
            ppos = pos; prem = rem;
        }
        cut[t] = idx;
    }
    for (unsigned t = 1; t <= T; ++t) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    std::vector<RunList> lists(T);
    std::vector<Stats> lstats(T);
    if (T == 1) sweep_range(cut[0], cut[1], lists[0], lstats[0]);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t]() { sweep_range(cut[t], cut[t + 1], lists[t], lstats[t]); });
        for (auto& x : th) x.join();
    }
    for (unsigned t = 0; t < T; ++t) {
        const RunList& rl = lists[t];
        for (size_t k = 0; k < rl.c.size(); ++k) rb.emit(rl.c[k], rl.cnt[k], rl.sf[k], rl.sl[k]);
        st.groups += lstats[t].groups; st.uniform_rows += lstats[t].uniform_rows; st.merged_rows += lstats[t].merged_rows;
    }
    if (rb.rows_emitted() != n) throw std::logic_error("prefix-free parsing: row count differs from the text length");
    lap("sweep");
    rb.finish();
    lap("finish (sort run-first samples)");
    if (stats) *stats = st;
    return L;
}

}  // namespace pfp

// Builder selection: prefix-free parsing from 16 MB up (identical result, memory ~ dictionary + parse; on a
// text that is not repetitive the dictionary approaches the text and a 2^31 limit may be hit: the in-memory
// SA-IS route is the fallback), SA-IS below. RIB_BUILDER=sais|pfp forces one.
inline LogicalIndex build_logical_index_auto(const uint8_t* text, uint64_t len, bool* used_pfp = nullptr) {
    const char* force = getenv("RIB_BUILDER");
    bool try_pfp = len >= ((uint64_t)1 << 24);
    if (force && std::strcmp(force, "sais") == 0) try_pfp = false;
    if (force && std::strcmp(force, "pfp") == 0) try_pfp = true;
    if (used_pfp) *used_pfp = false;
    if (try_pfp) {
        try {
            LogicalIndex L = pfp::build(text, len);
            if (used_pfp) *used_pfp = true;
            return L;
        } catch (const std::length_error&) {
        }
    }
    return build_logical_index(text, len);
}

}  // namespace rib
