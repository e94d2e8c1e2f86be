// cli_common.hpp — shared by ri-count / ri-locate: pattern-file reading, the reference's progress
// lines, and the multi-GPU fan-out (index replicated on every GPU, patterns cut into contiguous
// shards, one host thread per device, no collective on the search path — SURVEY.md §8e).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>
#include "logical_index.hpp"
#include "utils.hpp"
#include "../../include/rindex_gpu.h"

namespace ri {

struct PatternFile {
    uint64_t n = 0, m = 0;
    std::vector<uint8_t> body;  // n*m bytes
};

// Reference reader: one header line, then n*m raw bytes read with ifs.get (binary-safe: patterns may
// contain '\n' and bytes >= 0x80) — ri-count.cpp:86-110. A short file leaves the missing bytes 0.
inline PatternFile read_patterns(const std::string& path) {
    PatternFile pf;
    std::ifstream ifs(path, std::ios::binary);
    std::string header;
    std::getline(ifs, header);
    pf.n = get_number_of_patterns(header);
    pf.m = get_patterns_length(header);
    pf.body.assign(pf.n * pf.m, 0);
    ifs.read((char*)pf.body.data(), (std::streamsize)pf.body.size());
    return pf;
}

// The "% done ..." lines of the reference's loop (ri-count.cpp:96-102), in the same order.
inline void print_progress_lines(uint64_t n) {
    unsigned last_perc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        unsigned perc = (unsigned)((100 * i) / n);
        if (perc > last_perc) {
            std::cout << perc << "% done ..." << std::endl;
            last_perc = perc;
        }
    }
}

inline void die(int rc, const char* where) {
    std::cout << "Error: " << where << ": " << rig_strerror(rc);
    const char* ce = rig_last_cuda_error();
    if (ce && *ce) std::cout << " [" << ce << "]";
    std::cout << std::endl;
    exit(1);
}

// One rig_index per device; shard p covers patterns [N*p/G, N*(p+1)/G).
class GpuFleet {
public:
    // count_only: ri-count never expands occurrences, so the Phi tables are built at their smallest (D = 1, no seed table).
    // flat: file of the FLATTENED index (--flat): loaded instead of flattening when it exists and belongs to L
    // (rig_index_load_flat), written after the first flatten otherwise; such a file always holds the full tables.
    explicit GpuFleet(const rib::LogicalIndex& L, int gpus, bool count_only = false, const std::string& flat = std::string()) {
        int have = rig_device_count();
        if (have < 1) { std::cout << "Error: no CUDA device available (this build has no CPU query path)" << std::endl; exit(1); }
        G = gpus <= 0 ? 1 : (gpus > have ? have : gpus);
        idx.assign(G, nullptr);
        rig_logical_view v;
        v.n = L.n; v.r = L.r; v.F = L.F;
        v.run_heads = L.run_heads.data(); v.run_lens = L.run_lens.data(); v.samples_last = L.samples_last.data();
        v.pred_pos = L.pred_pos.data(); v.pred_to_run = L.pred_to_run.data();
        std::vector<int> rcs(G, 0), loaded(G, 0);
        rig_options opt;
        std::memset(&opt, 0, sizeof(opt));
        if (count_only && flat.empty()) { opt.reserved[0] = 1; opt.reserved[2] = 1; }
        run([&](int g) {
            if (!flat.empty()) {
                int rc = rig_index_load_flat(flat.c_str(), &v, g, &idx[g]);
                if (rc == RIG_OK) { loaded[g] = 1; return; }
                if (rc != RIG_ERR_INDEX && rc != RIG_ERR_ARG) { rcs[g] = rc; return; }   // absent / another index's file: flatten
            }
            rcs[g] = rig_index_create_ex(&v, g, &opt, &idx[g]);
        });
        for (int g = 0; g < G; ++g) if (rcs[g] != RIG_OK) die(rcs[g], "rig_index_create");
        if (!flat.empty() && !loaded[0]) {
            int rc = rig_index_save_flat(idx[0], flat.c_str());
            if (rc != RIG_OK) std::cout << "Warning: could not write the flattened index to " << flat << std::endl;
        }
        from_flat = !flat.empty() && loaded[0];
    }
    bool from_flat = false;
    ~GpuFleet() { for (auto* p : idx) rig_index_destroy(p); }
    int size() const { return G; }
    void attach_text(const uint8_t* text, uint64_t len) {  // -c: the indexed text goes to every GPU's HBM
        std::vector<int> rcs(G, 0);
        run([&](int g) { rcs[g] = rig_text_attach(idx[g], text, len); });
        for (int g = 0; g < G; ++g) {
            if (rcs[g] == RIG_ERR_ARG) { std::cout << "Error: the text given with -c is not the indexed text (length mismatch)" << std::endl; exit(0); }
            if (rcs[g] != RIG_OK) die(rcs[g], "rig_text_attach");
        }
    }
    rig_index* handle(int g) { return idx[g]; }

    void count(const uint8_t* patt, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi) {
        std::vector<int> rcs(G, 0);
        run([&](int g) {
            uint64_t a = N * g / G, b = N * (g + 1) / G;
            rcs[g] = rig_count_batch(idx[g], patt + a * m, b - a, m, lo + a, hi + a);
        });
        for (int g = 0; g < G; ++g) if (rcs[g] != RIG_OK) die(rcs[g], "rig_count_batch");
    }
    // SURVEY §8e re-balancing: contiguous shards of near-equal WORK, work(p) = n_occ(p) + 64 (the backward search of a
    // pattern costs about as much as a few dozen occurrences of expansion). Same rule as r-index_b200/_shard.py.
    static std::vector<uint64_t> balanced_cuts(const uint64_t* lo, const uint64_t* hi, uint64_t N, int G) {
        // integer arithmetic only: the rule of _shard.py / balanced_cuts_kernel, bit for bit
        std::vector<uint64_t> cum(N);
        uint64_t acc = 0;
        for (uint64_t p = 0; p < N; ++p) { acc += (hi[p] >= lo[p] ? hi[p] - lo[p] + 1 : 0) + 64; cum[p] = acc; }
        const unsigned __int128 W = (unsigned)G, total = acc;
        std::vector<uint64_t> cuts(1, 0);
        for (int k = 1; k < G; ++k) {
            uint64_t c = N;
            if (N && acc) {
                const unsigned __int128 t = total * (unsigned)k;
                const uint64_t need = (uint64_t)((t + W - 1) / W);   // first i with cum[i] >= ceil(t / G)
                const uint64_t i = uint64_t(std::lower_bound(cum.begin(), cum.end(), need) - cum.begin());
                if (i < N) {
                    const unsigned __int128 ci = cum[i], cp = i ? cum[i - 1] : 0;
                    c = (ci * W - t) > (t - cp * W) ? i : i + 1;
                }
            }
            cuts.push_back(std::min<uint64_t>(std::max<uint64_t>(c, cuts.back()), N));
        }
        cuts.push_back(N);
        return cuts;
    }

    // Locate the batch; the occurrences STAY ON THE DEVICES (one rig_locate_batch_ex call per shard with
    // RIG_LOCATE_DEVICE_ONLY: no second search, no download). With more than one GPU the shards are first counted
    // (equal-count shards), then re-cut at equal occurrence mass: cuts[g] .. cuts[g+1] are shard g's patterns.
    // off[g] (shard-local, size = shard patterns + 1) indexes shard g's occurrences; fetch(g, ...) copies them out.
    // flags: RIG_LOCATE_SORT (-o) / RIG_LOCATE_CHECK (-c, needs attach_text) run on the device
    // (ri-locate.cpp:146-190); reports[g] receives shard g's check report.
    uint64_t locate(const uint8_t* patt, uint64_t N, uint64_t m, uint64_t* lo, uint64_t* hi,
                    std::vector<std::vector<uint64_t>>& off, std::vector<uint64_t>& cuts, std::vector<uint64_t>& totals,
                    uint32_t flags = 0, std::vector<rig_check_report>* reports = nullptr) {
        off.assign(G, {});
        totals.assign(G, 0);
        if (G > 1) { count(patt, N, m, lo, hi); cuts = balanced_cuts(lo, hi, N, G); }
        else cuts = {0, N};
        std::vector<int> rcs(G, 0);
        std::vector<std::string> errs(G);
        std::vector<rig_check_report> reps(G);
        run([&](int g) {
            const uint64_t a = cuts[g], b = cuts[g + 1];
            off[g].assign(b - a + 1, 0);
            rcs[g] = rig_locate_batch_ex(idx[g], patt + a * m, b - a, m, lo + a, hi + a, off[g].data(), nullptr, 0, &totals[g],
                                         flags | RIG_LOCATE_DEVICE_ONLY, &reps[g]);
            if (rcs[g] != RIG_OK) errs[g] = rig_last_cuda_error();  // the CUDA error text is per host thread
        });
        if (reports) *reports = reps;
        uint64_t total = 0;
        for (int g = 0; g < G; ++g) {
            if (rcs[g] != RIG_OK) {
                std::cout << "Error: rig_locate_batch_ex: " << rig_strerror(rcs[g]);
                if (!errs[g].empty()) std::cout << " [" << errs[g] << "]";
                std::cout << std::endl;
                exit(1);
            }
            total += totals[g];
        }
        return total;
    }
    // shard g's occurrences [first, first + count) -> host (page-locked buffers from rig_host_alloc download fastest)
    void fetch(int g, uint64_t first, uint64_t count, uint64_t* out) {
        int rc = rig_fetch_occurrences(idx[g], first, count, out);
        if (rc != RIG_OK) die(rc, "rig_fetch_occurrences");
    }

private:
    template <class Fn>
    void run(Fn fn) {
        if (G == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back([&, g]() { fn(g); });
        for (auto& t : th) t.join();
    }
    int G = 1;
    std::vector<rig_index*> idx;
};

}  // namespace ri
