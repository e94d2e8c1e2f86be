// ri-build — builds the r-index of a text file (reference ri-build.cpp: same usage text, options,
// stdout lines and the leading 1-byte `fast` flag in the .ri file, :133). Construction is host-only:
// prefix-free parsing for texts from 16 MB up (host/pfp_builder.hpp), this repo's own SA-IS +
// run/sample scan below that (host/logical_index.hpp); both give the same arrays. The container is this repo's own.
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include "r_index.hpp"
#include "utils.hpp"

using namespace ri;
using namespace std;

static string out_basename, input_file;
static bool sais = true;

static void help() {
    cout << "ri-build: builds the r-index. Extension .ri is automatically added to output index file" << endl << endl;
    cout << "Usage: ri-build [options] <input_file_name>" << endl;
    cout << "   -o <basename>        use 'basename' as prefix for all index files. Default: basename is the specified input_file_name" << endl;
    cout << "   -divsufsort          use divsufsort algorithm to build the BWT (fast, 7.5n Bytes of RAM). By default," << endl;
    cout << "                        SE-SAIS is used (about 4 time slower than divsufsort, 4n Bytes of RAM)." << endl;
    cout << "   <input_file_name>    input text file." << endl;
    exit(0);
}

static void parse_args(char** argv, int argc, int& ptr) {
    string s(argv[ptr]);
    ptr++;
    if (s.compare("-o") == 0) {
        if (ptr >= argc - 1) {
            cout << "Error: missing parameter after -o option." << endl;
            help();
        }
        out_basename = string(argv[ptr]);
        ptr++;
    } else if (s.compare("-divsufsort") == 0) {
        sais = false;
    } else {
        cout << "Error: unrecognized '" << s << "' option." << endl;
        help();
    }
}

int main(int argc, char** argv) {
    using std::chrono::high_resolution_clock;
    auto t1 = high_resolution_clock::now();
    int ptr = 1;
    if (argc < 2) help();
    while (ptr < argc - 1) parse_args(argv, argc, ptr);
    input_file = string(argv[ptr]);
    if (out_basename.compare("") == 0) out_basename = string(input_file);
    string idx_file = out_basename;
    idx_file.append(".ri");
    cout << "Building r-index of input file " << input_file << endl;
    cout << "Index will be saved to " << idx_file << endl;
    string input;
    {
        std::ifstream fs(input_file, std::ios::binary);
        std::stringstream buffer;
        buffer << fs.rdbuf();
        input = buffer.str();
    }
    std::ofstream out(idx_file, std::ios::binary);
    bool fast = false;  // flag storing whether index is fast or small (reference ri-build.cpp:133)
    out.write((char*)&fast, sizeof(fast));
    {
        r_index<> idx(input, sais);
        idx.serialize(out);
    }
    auto t2 = high_resolution_clock::now();
    uint64_t total = std::chrono::duration_cast<std::chrono::duration<double, std::ratio<1>>>(t2 - t1).count();
    cout << "Build time : " << get_time(total) << endl;
    out.close();
}
