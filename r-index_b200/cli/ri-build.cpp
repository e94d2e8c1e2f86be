// ri-build — builds the r-index of a text file (reference ri-build.cpp: same usage text, options,
// stdout lines and the leading 1-byte `fast` flag in the .ri file, :133). Construction is host-only:
// prefix-free parsing for texts from 16 MB up (host/pfp_builder.hpp), this repo's own SA-IS +
// run/sample scan below that (host/logical_index.hpp); both give the same arrays. The container is this repo's own.
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
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
    // The reference copies the file into a string (ri-build.cpp:121-129). Here the file is mapped: the prefix-free
    // parsing builder scans the text front to back twice and otherwise touches it at the offsets of its dictionary
    // phrases (comparisons, sorting, the dictionary fill), keeping only (offset, length) references into it, so the
    // resident memory of a large build is the dictionary and the parse, not the text. No access-pattern hint is
    // given: MADV_SEQUENTIAL would drop pages the phrase comparisons come back to. The file must not be truncated
    // while the build runs (a mapped read past the new end raises SIGBUS, which a private copy would not).
    int fd = open(input_file.c_str(), O_RDONLY);
    struct stat sb;
    const uint8_t* text = nullptr;
    size_t text_len = 0;
    string fallback;
    void* mapped = nullptr;
    if (fd >= 0 && fstat(fd, &sb) == 0 && sb.st_size > 0) {
        void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) { text = (const uint8_t*)m; text_len = (size_t)sb.st_size; mapped = m; }
    }
    if (!text) {  // empty or unmappable input (a pipe, ...): read it as the reference does
        std::ifstream fs(input_file, std::ios::binary);
        std::stringstream buffer;
        buffer << fs.rdbuf();
        fallback = buffer.str();
        text = (const uint8_t*)fallback.data(); text_len = fallback.size();
    }
    std::ofstream out(idx_file, std::ios::binary);
    bool fast = false;  // flag storing whether index is fast or small (reference ri-build.cpp:133)
    out.write((char*)&fast, sizeof(fast));
    {
        r_index<> idx(text, text_len, sais);
        idx.serialize(out);
    }
    if (mapped) munmap(mapped, text_len);
    if (fd >= 0) close(fd);
    auto t2 = high_resolution_clock::now();
    uint64_t total = std::chrono::duration_cast<std::chrono::duration<double, std::ratio<1>>>(t2 - t1).count();
    cout << "Build time : " << get_time(total) << endl;
    out.close();
}
