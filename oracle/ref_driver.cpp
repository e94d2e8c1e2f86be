// ref_driver.cpp — C entry points over the UNMODIFIED reference headers (oracle/_ref/libri_ref.so).
//
// TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl
// reference). Never linked or loaded by the product path.
//
// The reference sources are compiled where they lie (-I/root/reference -I/root/reference/internal,
// see oracle/Makefile); sdsl-lite is replaced by oracle/sdsl_shim. What runs here
// is therefore the reference's own r_index<>::count / locate_all / Phi / rle_string::rank code
// (internal/r_index.hpp:171-221,292-355,482-545; internal/rle_string.hpp:126-256), over restated
// SDSL primitives.
#include <string>
#include <vector>
#include <set>
#include <tuple>
#include <sstream>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <cstring>
#include <cmath>
#include <chrono>
#include <queue>
#include <map>
#include <memory>
#include <thread>
#include <atomic>
#include <streambuf>
#define private public  // read-only access to r_index<> members for ref_extract (test infra only)
#include "internal/r_index.hpp"
#undef private
#include <thread>
#include <atomic>
#include <streambuf>

using namespace ri;
typedef r_index<> ref_index_t;

namespace {
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
struct Quiet {  // the reference ctor prints progress to cout (r_index.hpp:53-57); silence it on request
    std::streambuf* old = nullptr; NullBuf nb;
    explicit Quiet(bool on) { if (on) old = std::cout.rdbuf(&nb); }
    ~Quiet() { if (old) std::cout.rdbuf(old); }
};
template <class Fn>
void shard(uint64_t N, int nthreads, Fn fn) {
    if (nthreads <= 1) { fn(0, N, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        uint64_t a = N * (uint64_t)t / (uint64_t)nthreads, b = N * (uint64_t)(t + 1) / (uint64_t)nthreads;
        th.emplace_back([=]() { fn(a, b, t); });
    }
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

// Returns NULL if the text holds 0x00/0x01 (the reference would exit(1), r_index.hpp:46-51).
void* ref_build(const uint8_t* text, uint64_t len, int quiet) {
    for (uint64_t i = 0; i < len; ++i) if (text[i] == 0 || text[i] == 1) return nullptr;
    Quiet q(quiet != 0);
    std::string s((const char*)text, len);
    return new ref_index_t(s, true);
}
void ref_free(void* h) { delete (ref_index_t*)h; }
uint64_t ref_bwt_size(void* h) { return ((ref_index_t*)h)->bwt_size(); }
uint64_t ref_number_of_runs(void* h) { return ((ref_index_t*)h)->number_of_runs(); }

int ref_save(void* h, const char* path) {  // same framing as ri-build.cpp:129-144
    std::ofstream out(path);
    if (!out) return -1;
    bool fast = false;
    out.write((char*)&fast, sizeof(fast));
    ((ref_index_t*)h)->serialize(out);
    return out ? 0 : -1;
}
void* ref_load(const char* path) {  // same framing as ri-count.cpp:153-158
    std::ifstream in(path);
    if (!in) return nullptr;
    bool fast;
    in.read((char*)&fast, sizeof(fast));
    ref_index_t* idx = new ref_index_t();
    idx->load(in);
    return idx;
}

// N calls of r_index<>::count. Returns seconds spent in the query loop.
double ref_count_batch(void* h, const uint8_t* patt, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi, int nthreads) {
    ref_index_t* idx = (ref_index_t*)h;
    auto t0 = std::chrono::steady_clock::now();
    shard(N, nthreads, [&](uint64_t a, uint64_t b, int) {
        for (uint64_t p = a; p < b; ++p) {
            std::string P((const char*)patt + p * m, m);
            range_t rn = idx->count(P);
            if (lo) lo[p] = rn.first;
            if (hi) hi[p] = rn.second;
        }
    });
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// N calls of r_index<>::locate_all. occ_offsets[N+1] must already hold the exclusive prefix sums
// of the per-pattern counts when occ != NULL (use ref_count_batch first); with occ == NULL the
// vectors are dropped as ri-locate does (ri-locate.cpp:144). Returns seconds in the query loop;
// *occ_total = number of occurrences produced.
double ref_locate_batch(void* h, const uint8_t* patt, uint64_t N, uint64_t m, const uint64_t* occ_offsets,
                        uint64_t* occ, uint64_t* occ_total, int nthreads) {
    ref_index_t* idx = (ref_index_t*)h;
    std::atomic<uint64_t> total(0);
    auto t0 = std::chrono::steady_clock::now();
    shard(N, nthreads, [&](uint64_t a, uint64_t b, int) {
        uint64_t mine = 0;
        for (uint64_t p = a; p < b; ++p) {
            std::string P((const char*)patt + p * m, m);
            std::vector<ulint> OCC = idx->locate_all(P);
            mine += OCC.size();
            if (occ) std::memcpy(occ + occ_offsets[p], OCC.data(), OCC.size() * sizeof(ulint));
        }
        total += mine;
    });
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (occ_total) *occ_total = total.load();
    return dt;
}

// Single-operation probes for unit parity of the device primitives.
uint64_t ref_bwt_rank(void* h, uint64_t i, uint8_t c) { return ((ref_index_t*)h)->bwt.rank(i, c); }
uint8_t ref_bwt_at(void* h, uint64_t i) { return ((ref_index_t*)h)->bwt[i]; }
uint64_t ref_phi(void* h, uint64_t i) { return ((ref_index_t*)h)->Phi(i); }
// single-position navigation through the reference's own methods (r_index.hpp:162-164, 224-271)
//   op 0: operator[](i)   1: LF(i)   2: FL(i)   3: F_at(i)
void ref_navigate(void* h, int op, const uint64_t* pos, uint64_t N, uint64_t* out) {
    ref_index_t* idx = (ref_index_t*)h;
    for (uint64_t k = 0; k < N; ++k) {
        const uint64_t i = pos[k];
        out[k] = op == 0 ? (uint64_t)(*idx)[i] : op == 1 ? idx->LF(i) : op == 2 ? idx->FL(i) : (uint64_t)idx->F_at(i);
    }
}
// rle_string::break_range (rle_string.hpp:261-302) through the reference's own method. Two-call use: with first == NULL
// only the number of sub-ranges is returned.
uint64_t ref_break_range(void* h, uint64_t l, uint64_t r, uint8_t c, uint64_t* first, uint64_t* last) {
    auto v = ((ref_index_t*)h)->bwt.break_range({l, r}, c);
    if (first) for (size_t k = 0; k < v.size(); ++k) { first[k] = v[k].first; last[k] = v[k].second; }
    return v.size();
}
// rle_string::closest_run_break (rle_string.hpp:455-493)
uint64_t ref_closest_run_break(void* h, uint64_t l, uint64_t r, uint8_t c) { return ((ref_index_t*)h)->bwt.closest_run_break({l, r}, c); }
// the BWT symbol by symbol through operator[] (what r_index::get_bwt / rle_string::toString return, r_index.hpp:375-377)
void ref_get_bwt(void* h, uint8_t* out) {
    ref_index_t* idx = (ref_index_t*)h;
    const uint64_t n = idx->bwt_size();
    for (uint64_t i = 0; i < n; ++i) out[i] = (*idx)[i];
}

// A reference r_index<> assembled from the LOGICAL content of an index (the arrays ref_extract returns), by the
// reference's own structure constructors: rle_string(string&) on the BWT spelled out from its runs (rle_string.hpp:
// 52-124) and the statements of the r_index constructor that follow the suffix sort (r_index.hpp:67-146: F, the
// terminator position, pred, samples_last, pred_to_run with the reference's bit widths). Only sufsort() (:553-634)
// is bypassed: at 4 GB its suffix array needs half an hour, the structures below a minute. Used by bench.py's
// reference arm when no reference-built .ri of the workload travelled to the box; the logical arrays come from this
// repo's builder, whose output equals sufsort()'s on every text where both have been run (tests/test_host.py).
void* ref_from_logical(uint64_t n, uint64_t r, const uint8_t* heads, const uint64_t* lens, const uint64_t* samples_last,
                       const uint64_t* pred_pos, const uint64_t* pred_to_run) {
    std::string bwt_s;
    bwt_s.reserve(n);
    for (uint64_t j = 0; j < r; ++j) bwt_s.append(lens[j], (char)heads[j]);
    if (bwt_s.size() != n) return nullptr;
    ref_index_t* idx = new ref_index_t();
    idx->bwt = rle_string_sd(bwt_s);                                        // r_index.hpp:67
    idx->F = std::vector<ulint>(256, 0);                                    // :70-82
    for (uchar c : bwt_s) idx->F[c]++;
    for (ulint i = 255; i > 0; --i) idx->F[i] = idx->F[i - 1];
    idx->F[0] = 0;
    for (ulint i = 1; i < 256; ++i) idx->F[i] += idx->F[i - 1];
    for (ulint i = 0; i < bwt_s.size(); ++i)                                // :84-86
        if (bwt_s[i] == ref_index_t::TERMINATOR) idx->terminator_position = i;
    idx->r = idx->bwt.number_of_runs();                                     // :92
    if (idx->r != r) { delete idx; return nullptr; }
    const int log_r = bitsize(uint64_t(r)), log_n = bitsize(uint64_t(idx->bwt.size()));  // :97-98
    {
        auto pred_bv = std::vector<bool>(n, false);                         // :112-123
        for (uint64_t k = 0; k < r; ++k) pred_bv[pred_pos[k]] = true;
        idx->pred = sparse_sd_vector(pred_bv);
    }
    std::string().swap(bwt_s);
    idx->samples_last = int_vector<>(r, 0, log_n);                          // :131-146
    idx->pred_to_run = int_vector<>(r, 0, log_r);
    for (uint64_t i = 0; i < r; ++i) { idx->samples_last[i] = samples_last[i]; idx->pred_to_run[i] = pred_to_run[i]; }
    return idx;
}

// The logical content of a reference-built/loaded index, read through the reference's own
// accessors. This is also the extraction INTEGRATION.md proposes for feeding rig_index_create.
// Arrays: F[257], heads[r], lens[r], samples_last[r], pred_pos[r], pred_to_run[r].
void ref_extract(void* h, uint64_t* F, uint8_t* heads, uint64_t* lens, uint64_t* samples_last, uint64_t* pred_pos,
                 uint64_t* pred_to_run) {
    ref_index_t* idx = (ref_index_t*)h;
    uint64_t r = idx->number_of_runs();
    for (int c = 0; c < 256; ++c) F[c] = idx->F[c];
    F[256] = idx->bwt_size();
    for (uint64_t j = 0; j < r; ++j) {
        heads[j] = idx->bwt.run_heads[j];        // rle_string.hpp:568
        lens[j] = idx->bwt.run_at(j);            // rle_string.hpp:331-338
        samples_last[j] = idx->samples_last[j];  // r_index.hpp:664
        pred_pos[j] = idx->pred.select(j);       // r_index.hpp:663, sparse_sd_vector.hpp:178
        pred_to_run[j] = idx->pred_to_run[j];    // r_index.hpp:665
    }
}

}  // extern "C"
