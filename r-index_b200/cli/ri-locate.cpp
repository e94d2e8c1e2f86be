// ri-locate — locate all occurrences of the input patterns (reference ri-locate.cpp: same usage text,
// options -c / -o, argument handling and stdout lines). The per-pattern loop (:126-192) becomes one
// batch call per GPU shard. -o keeps the reference's format: per pattern, positions sorted ascending,
// printed through an (int) cast (:146-152; positions >= 2^31 therefore wrap exactly as upstream).
// The per-pattern sort of -o/-c and the -c self-check (:156-190) run on the GPU (rig_locate_batch_ex:
// segmented sort; brute-force counts as a hash join over the text, byte comparison of every occurrence).
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include "cli_common.hpp"

using namespace ri;
using namespace std;

static string check_text_file, ofile;
static int gpus = 1;
static string flat_file;

static void help() {
    cout << "ri-locate: locate all occurrences of the input patterns." << endl << endl;
    cout << "Usage: ri-locate [options] <index> <patterns>" << endl;
    cout << "   -c <text>    check correctness of each pattern occurrence on this text file (must be the same indexed)" << endl;
    cout << "   -o <ofile>   write pattern occurrences to this file (ASCII)" << endl;
    cout << "   <index>      index file (with extension .ri)" << endl;
    cout << "   <patterns>   file in pizza&chili format containing the patterns." << endl;
    exit(0);
}

static void parse_args(char** argv, int argc, int& ptr) {
    string s(argv[ptr]);
    ptr++;
    if (s.compare("-c") == 0) {
        if (ptr >= argc - 1) {
            cout << "Error: missing parameter after -c option." << endl;
            help();
        }
        check_text_file = string(argv[ptr]);
        ptr++;
    } else if (s.compare("-o") == 0) {
        if (ptr >= argc - 1) {
            cout << "Error: missing parameter after -o option." << endl;
            help();
        }
        ofile = string(argv[ptr]);
        ptr++;
    } else if (s.compare("--gpus") == 0 && ptr < argc - 2) {  // addition
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
    in.read((char*)&fast, sizeof(fast));
    cout << "Loading r-index" << endl;

    string text;
    bool c = false;
    ofstream out;
    if (ofile.compare(string()) != 0) out = ofstream(ofile);
    if (check_text_file.compare(string()) != 0) {
        c = true;
        ifstream ifs1(check_text_file, std::ios::binary);
        stringstream ss;
        ss << ifs1.rdbuf();
        text = ss.str();
    }

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
    GpuFleet fleet(L, gpus, false, flat_file);                    // flatten + upload: accounted as load time, not search time
    auto u2 = high_resolution_clock::now();
    const uint64_t n = pf.n, m = pf.m;
    std::vector<uint64_t> lo(n), hi(n);
    std::vector<std::vector<uint64_t>> off;
    std::vector<uint64_t> cuts, totals;
    uint32_t flags = 0;
    if (ofile.compare(string()) != 0) flags |= RIG_LOCATE_SORT;
    if (c) { flags |= RIG_LOCATE_CHECK; fleet.attach_text((const uint8_t*)text.data(), text.size()); }
    std::vector<rig_check_report> reports;
    // the occurrences stay on the devices: the reference drops each pattern's vector unless -o / -c is given
    // (ri-locate.cpp:144); the -c check itself runs on the device; only -o needs the positions on the host
    uint64_t occ_tot = fleet.locate(pf.body.data(), n, m, lo.data(), hi.data(), off, cuts, totals, flags, &reports);

    const int G = fleet.size();
    if (ofile.compare(string()) != 0) {
        const uint64_t CH = 1ull << 24;  // 128 MB of positions per download, page-locked
        uint64_t* buf = (uint64_t*)rig_host_alloc(CH * 8);
        if (!buf) die(RIG_ERR_NOMEM, "rig_host_alloc");
        for (int g = 0; g < G; ++g)      // already sorted per pattern on the device (reference :147)
            for (uint64_t f = 0; f < totals[g]; f += CH) {
                const uint64_t k = std::min<uint64_t>(CH, totals[g] - f);
                fleet.fetch(g, f, k, buf);
                for (uint64_t t = 0; t < k; ++t) out << (int)buf[t] << endl;
            }
        rig_host_free(buf);
    }
    if (c) {  // check occurrences (reference :156-190): the first offending pattern, in shard order
        for (int g = 0; g < G; ++g) {
            const rig_check_report& r = reports[g];
            const uint64_t a = cuts[g];
            if (r.wrong_count_patterns || r.unsorted_or_duplicate) {
                const uint64_t i = a + (r.first_bad_pattern != ~0ull ? r.first_bad_pattern : 0);
                const uint64_t want = hi[i] >= lo[i] ? (hi[i] - lo[i]) + 1 : 0;
                cout << "Error: wrong number of located occurrences for pattern " << i << " (located " << want << "; "
                     << r.wrong_count_patterns << " patterns disagree with the text, " << r.unsorted_or_duplicate
                     << " repeated positions)" << endl;
                exit(0);
            }
            if (r.wrong_occurrences)
                cout << "Error: wrong occurrence: " << r.first_bad_position << " (" << occ_tot << " occurrences" << ") " << endl;
        }
    }
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
    rig_timing t;
    if (rig_last_timing(fleet.handle(0), &t) == RIG_OK)
        cout << "[gpu] devices = " << G << ", device 0: search = " << t.search_ms << " ms, scan = " << t.scan_ms
             << " ms, phi expansion = " << t.expand_ms << " ms, chains = " << t.chains << endl;
    in.close();
}
