// ri-count — number of occurrences of the input patterns (reference ri-count.cpp: same usage text,
// argument handling and stdout lines). The per-pattern loop of the reference (:96-114) becomes one
// batch call into the CUDA library; extra flags are additions and do not change the reference's.
#include <chrono>
#include <fstream>
#include <iostream>
#include "cli_common.hpp"

using namespace ri;
using namespace std;

static int gpus = 1;
static string flat_file;

static void help() {
    cout << "ri-count: number of occurrences of the input patterns." << endl << endl;
    cout << "Usage: ri-count <index> <patterns>" << endl;
    cout << "   <index>      index file (with extension .ri)" << endl;
    cout << "   <patterns>   file in pizza&chili format containing the patterns." << endl;
    exit(0);
}

static void parse_args(char** argv, int argc, int& ptr) {
    string s(argv[ptr]);
    ptr++;
    if (s.compare("--gpus") == 0 && ptr < argc - 2) {  // addition: shard the patterns over N GPUs
        gpus = atoi(argv[ptr]);
        ptr++;
    } else if (s.compare("--flat") == 0 && ptr < argc - 2) {  // addition: file of the flattened index (loaded if present, else written)
        flat_file = string(argv[ptr]);
        ptr++;
    } else {
        cout << "Error: unknown option " << s << endl;
        help();
    }
}

int main(int argc, char** argv) {
    using std::chrono::high_resolution_clock;
    if (argc < 3) help();
    int ptr = 1;
    while (ptr < argc - 2) parse_args(argv, argc, ptr);
    string idx_file(argv[ptr]);
    string patt_file(argv[ptr + 1]);
    std::ifstream in(idx_file, std::ios::binary);
    bool fast;
    in.read((char*)&fast, sizeof(fast));  // fast or small index? (reference ri-count.cpp:155-158)
    cout << "Loading r-index" << endl;

    auto t1 = high_resolution_clock::now();
    rib::LogicalIndex L;
    if (!rib::load(L, in)) {
        cout << "Error: index file is not an r-index built by this ri-build" << endl;
        exit(1);
    }
    auto t2 = high_resolution_clock::now();
    cout << "searching patterns ... " << endl;
    PatternFile pf = read_patterns(patt_file);  // a malformed header exits(0) here, as upstream (utils.hpp:51-55)
    auto u1 = high_resolution_clock::now();
    GpuFleet fleet(L, gpus, true, flat_file);                    // flatten + upload: accounted as load time, not search time
    auto u2 = high_resolution_clock::now();
    const uint64_t n = pf.n, m = pf.m;
    std::vector<uint64_t> lo(n), hi(n);
    fleet.count(pf.body.data(), n, m, lo.data(), hi.data());
    uint64_t occ_tot = 0;
    for (uint64_t i = 0; i < n; ++i) occ_tot += hi[i] >= lo[i] ? (hi[i] - lo[i]) + 1 : 0;  // r_index::occ :307-313
    print_progress_lines(n);
    double occ_avg = (double)occ_tot / n;
    cout << endl << occ_avg << " average occurrences per pattern" << endl;
    auto t3 = high_resolution_clock::now();

    uint64_t upload = std::chrono::duration_cast<std::chrono::milliseconds>(u2 - u1).count();
    uint64_t load = std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count() + upload;
    cout << "Load time : " << load << " milliseconds" << endl;
    uint64_t search = std::chrono::duration_cast<std::chrono::milliseconds>(t3 - t2).count() - upload;
    cout << "number of patterns n = " << n << endl;
    cout << "pattern length m = " << m << endl;
    cout << "total number of occurrences  occ_t = " << occ_tot << endl;
    cout << "Total time : " << search << " milliseconds" << endl;
    cout << "Search time : " << (double)search / n << " milliseconds/pattern (total: " << n << " patterns)" << endl;
    cout << "Search time : " << (double)search / occ_tot << " milliseconds/occurrence (total: " << occ_tot << " occurrences)" << endl;
    // additions (after the reference's lines)
    rig_timing t;
    if (rig_last_timing(fleet.handle(0), &t) == RIG_OK)
        cout << "[gpu] devices = " << fleet.size() << ", search kernel (device 0) = " << t.search_ms << " ms, LF steps = "
             << t.lf_steps << endl;
    in.close();
}
