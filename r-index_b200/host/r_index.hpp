// r_index.hpp — host-side mirror of the reference's public index surface, backed by the CUDA library.
//
// Same names, argument meaning and error behaviour as ri::r_index<> (reference
// internal/r_index.hpp:28-667) for the members the count/locate path and the three CLIs use:
//   r_index()                      :37     r_index(string&, bool sais) :42   (build; same stdout lines)
//   range_t count(string&)         :292    ulint occ(string&)          :307
//   vector<ulint> locate_all(string&) :328 ulint serialize(ostream&)   :382  void load(istream&) :407
//   number_of_runs() :361  text_size() :450  bwt_size() :454  get_terminator_position() :368
//   operator[](i) :162  LF(i) :224  FL(i) :232  F_at(i) :263  get_char_range(c) :276  get_bwt() :375
//   print_space() :462   and, of the BWT member (rle_string.hpp): break_range(rn, c) :261, closest_run_break(rn, c) :455
// plus the batch surface this repo adds (one FFI call per batch instead of one call per pattern):
//   count_batch(), locate_batch(), navigate_batch().
// Every query — single-pattern calls included — runs on the GPU through include/rindex_gpu.h.
// There is no CPU query path: if no device is available the query members print the error and exit(1).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <string>
#include <vector>
#include <utility>
#include "logical_index.hpp"
#include "pfp_builder.hpp"
#include "../../include/rindex_gpu.h"

namespace ri {

typedef uint64_t ulint;
typedef unsigned char uchar;
typedef std::pair<ulint, ulint> range_t;

inline uint8_t bitsize(uint64_t x) {  // reference internal/utils.hpp:43-48
    if (x == 0) return 1;
    return 64 - __builtin_clzll(x);
}

template <class sparse_bv_type = void, class rle_string_t = void>  // template kept so that `r_index<>` compiles
class r_index {
public:
    r_index() {}
    ~r_index() { release_device(); }
    r_index(const r_index&) = delete;
    r_index& operator=(const r_index&) = delete;

    // Build index (reference r_index.hpp:42-150: same progress lines on stdout, exit(1) on reserved bytes).
    r_index(std::string& input, bool sais = true) : r_index((const uint8_t*)input.data(), input.size(), sais) {}
    // the same constructor over a byte range (ri-build maps large files instead of copying them into a string)
    r_index(const uint8_t* text, size_t text_len, bool sais = true) {
        using std::cout; using std::endl; using std::flush;
        struct View { const uint8_t* p; size_t n; const uint8_t* data() const { return p; } size_t size() const { return n; } } input{text, text_len};
        if (rib::contains_reserved_chars((const uint8_t*)input.data(), input.size())) {
            cout << "Error: input string contains one of the reserved characters 0x0, 0x1" << endl;
            exit(1);
        }
        cout << "Text length = " << input.size() << endl << endl;
        cout << "(1/3) Building BWT and computing SA samples";
        if (sais) cout << " (SE-SAIS) ... " << flush;
        else cout << "(DIVSUFSORT) ... " << flush;  // both flags use this repo's own in-memory SA-IS
        // construction: prefix-free parsing for large texts, in-memory SA-IS otherwise (pfp_builder.hpp; same arrays)
        L = rib::build_logical_index_auto((const uint8_t*)input.data(), input.size());
        cout << "done.\n(2/3) RLE encoding BWT ... " << flush;
        cout << "done. " << endl << endl;
        cout << "Number of BWT equal-letter runs: r = " << L.r << endl;
        cout << "Rate n/r = " << double(L.n) / L.r << endl;
        cout << "log2(r) = " << log2(double(L.r)) << endl;
        cout << "log2(n/r) = " << log2(double(L.n) / L.r) << endl << endl;
        cout << "(3/3) Building phi function ..." << flush;
        cout << " done. " << endl << endl;
    }

    range_t full_range() { return {0, bwt_size() - 1}; }

    // Return BWT range of pattern P (reference :292-302). Empty range = {1,0}.
    range_t count(std::string& P) {
        ulint lo = 0, hi = 0;
        count_batch((const uint8_t*)P.data(), 1, P.size(), &lo, &hi);
        return {lo, hi};
    }
    // Number of occurrences of P (reference :307-313).
    ulint occ(std::string& P) {
        auto rn = count(P);
        return rn.second >= rn.first ? (rn.second - rn.first) + 1 : 0;
    }
    // All occurrences of P, in the reference's order SA[hi], SA[hi-1], ..., SA[lo] (reference :328-355).
    std::vector<ulint> locate_all(std::string& P) {
        ulint lo = 0, hi = 0;
        std::vector<ulint> off, occv;
        locate_batch((const uint8_t*)P.data(), 1, P.size(), &lo, &hi, off, occv);
        return occv;
    }

    // ---- batch surface: N patterns of fixed length m, contiguous (the Pizza&Chili body) ----
    void count_batch(const uint8_t* patt, ulint N, ulint m, ulint* lo, ulint* hi) {
        ensure_device();
        check(rig_count_batch(dev, patt, N, m, lo, hi), "rig_count_batch");
    }
    // occ_offsets gets N+1 entries; returns the total number of occurrences. One search: the occurrences are located
    // into the library's device buffer (RIG_LOCATE_DEVICE_ONLY), then downloaded once their number is known.
    ulint locate_batch(const uint8_t* patt, ulint N, ulint m, ulint* lo, ulint* hi, std::vector<ulint>& occ_offsets,
                       std::vector<ulint>& occ) {
        ensure_device();
        occ_offsets.assign(N + 1, 0);
        uint64_t total = 0;
        check(rig_locate_batch_ex(dev, patt, N, m, lo, hi, occ_offsets.data(), nullptr, 0, &total, RIG_LOCATE_DEVICE_ONLY, nullptr),
              "rig_locate_batch_ex");
        occ.resize(total);
        check(rig_fetch_occurrences(dev, 0, total, occ.data()), "rig_fetch_occurrences");
        return total;
    }

    // ---- single-position navigation (reference :162-164, :224-271, :375-377), each a batch of one on the device;
    //      navigate_batch / get_bwt take whole arrays ----
    uchar operator[](ulint i) { return (uchar)navigate1(RIG_NAV_BWT, i); }
    ulint LF(ulint i) { return navigate1(RIG_NAV_LF, i); }
    ulint FL(ulint i) { return navigate1(RIG_NAV_FL, i); }
    uchar F_at(ulint i) { return (uchar)navigate1(RIG_NAV_F_AT, i); }
    void navigate_batch(int op, const ulint* positions, ulint N, ulint* out) {
        ensure_device();
        check(rig_navigate_batch(dev, op, positions, N, out), "rig_navigate_batch");
    }
    std::string get_bwt() {  // the BWT as a string, terminator row = 0x01 (rle_string::toString)
        ensure_device();
        std::string s(L.n, '\0');
        check(rig_get_bwt(dev, 0, L.n, (uint8_t*)&s[0]), "rig_get_bwt");
        return s;
    }
    // BWT range of character c (reference :276-290): {1,0} if c does not occur
    range_t get_char_range(uchar c) {
        if (L.F[c] >= L.F[(unsigned)c + 1]) return {1, 0};
        return {L.F[c], L.F[(unsigned)c + 1] - 1};
    }

    // rle_string::break_range (reference rle_string.hpp:261-302): maximal sub-ranges of rn holding only c; requires
    // bwt[rn.first] == bwt[rn.second] == c (the reference asserts; here a violating call returns no range)
    std::vector<range_t> break_range(range_t rn, uchar c) {
        ensure_device();
        uint64_t off[2] = {0, 0}, total = 0;
        int rc = rig_break_range_batch(dev, &rn.first, &rn.second, &c, 1, off, nullptr, nullptr, 0, &total);
        std::vector<ulint> a(total), b(total);
        if (rc == RIG_ERR_CAPACITY) rc = rig_break_range_batch(dev, &rn.first, &rn.second, &c, 1, off, a.data(), b.data(), total, &total);
        check(rc, "rig_break_range_batch");
        std::vector<range_t> out(total);
        for (uint64_t k = 0; k < total; ++k) out[k] = {a[k], b[k]};
        return out;
    }
    // rle_string::closest_run_break (reference rle_string.hpp:455-493)
    ulint closest_run_break(range_t rn, uchar c) {
        ensure_device();
        ulint out = 0;
        check(rig_closest_run_break_batch(dev, &rn.first, &rn.second, &c, 1, &out), "rig_closest_run_break_batch");
        return out;
    }
    // r_index::print_space (reference :462-472 over rle_string::print_space, rle_string.hpp:402-442): the same report
    // lines. The byte counts are those of THIS index's containers — the run-length BWT as plain arrays on the host and
    // its flattened form in HBM — not of SDSL's serialized structures, which this repo does not have.
    ulint print_space() {
        using std::cout; using std::endl;
        cout << "Number of runs = " << L.r << endl << endl;
        const ulint runs_bytes = L.run_lens.size() * sizeof(uint64_t), heads_bytes = L.run_heads.size();
        cout << "main runs bitvector: " << runs_bytes << " Bytes" << endl;          // run lengths (the role of `runs`)
        cout << "runs-per-letter bitvectors: " << 0 << " Bytes" << endl;             // folded into the block records on the device
        cout << "run heads: " << heads_bytes << " Bytes" << endl;
        const ulint tot_bytes = runs_bytes + heads_bytes;
        cout << "\nTOT BWT space: " << tot_bytes << " Bytes" << endl << endl;
        if (dev) cout << "flattened index in HBM (all tables): " << info.device_bytes << " Bytes" << endl << endl;
        return tot_bytes;
    }

    ulint number_of_runs() { return L.r; }
    ulint get_terminator_position() { return L.terminator_position; }
    ulint text_size() { return L.n - 1; }
    ulint bwt_size() { return L.n; }
    uchar get_terminator() { return rib::kTerminator; }

    // serialize / load (reference :382-422). Container = this repo's own (logical_index.hpp); byte
    // compatibility with SDSL-serialized .ri files is out of scope.
    ulint serialize(std::ostream& out) { return rib::serialize(L, out); }
    void load(std::istream& in) {
        release_device();
        if (!rib::load(L, in)) {
            std::cout << "Error: index file is not an r-index built by this ri-build" << std::endl;
            exit(1);
        }
    }

    // ---- device control (additions) ----
    void set_device(int d) { if (d != device) { release_device(); device = d; } }
    const rig_index_info& device_info() { ensure_device(); return info; }
    rig_timing last_timing() { rig_timing t; ensure_device(); rig_last_timing(dev, &t); return t; }
    rig_index* device_handle() { ensure_device(); return dev; }
    const rib::LogicalIndex& logical() const { return L; }

private:
    void ensure_device() {
        if (dev) return;
        rig_logical_view v;
        v.n = L.n; v.r = L.r; v.F = L.F;
        v.run_heads = L.run_heads.data(); v.run_lens = L.run_lens.data(); v.samples_last = L.samples_last.data();
        v.pred_pos = L.pred_pos.data(); v.pred_to_run = L.pred_to_run.data();
        check(rig_index_create(&v, device, &dev), "rig_index_create");
        rig_index_info_get(dev, &info);
    }
    void release_device() { if (dev) { rig_index_destroy(dev); dev = nullptr; } }
    ulint navigate1(int op, ulint i) {
        ulint out = 0;
        navigate_batch(op, &i, 1, &out);
        return out;
    }
    void check(int rc, const char* where) {
        if (rc == RIG_OK) return;
        std::cout << "Error: " << where << ": " << rig_strerror(rc);
        const char* ce = rig_last_cuda_error();
        if (ce && *ce) std::cout << " [" << ce << "]";
        std::cout << std::endl;
        exit(1);
    }

    rib::LogicalIndex L;
    rig_index* dev = nullptr;
    rig_index_info info{};
    int device = 0;
};

}  // namespace ri
