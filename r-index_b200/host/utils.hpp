// utils.hpp — CLI-side helpers with the behaviour of the reference's internal/utils.hpp.
//   get_time            :14-41   "N seconds. (Hh Mm Ss)" formatting used by ri-build
//   header_error        :51-55   message + exit(0) on a malformed Pizza&Chili header
//   get_number_of_patterns :57-73, get_patterns_length :75-91
//       the text after "number=" / "length=" up to the next space, converted with atoi
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace ri {

inline std::string get_time(uint64_t time) {
    std::stringstream ss;
    if (time >= 3600) {
        uint64_t h = time / 3600, m = (time % 3600) / 60, s = (time % 3600) % 60;
        ss << time << " seconds. (" << h << "h " << m << "m " << s << "s" << ")";
    } else if (time >= 60) {
        uint64_t m = time / 60, s = time % 60;
        ss << time << " seconds. (" << m << "m " << s << "s" << ")";
    } else {
        ss << time << " seconds.";
    }
    return ss.str();
}

inline void header_error() {
    std::cout << "Error: malformed header in patterns file" << std::endl;
    std::cout << "Take a look here for more info on the file format: http://pizzachili.dcc.uchile.cl/experiments.html" << std::endl;
    exit(0);
}

inline uint64_t header_field(const std::string& header, const char* key) {
    size_t start_pos = header.find(key);
    if (start_pos == std::string::npos || start_pos + 7 >= header.size()) header_error();
    start_pos += 7;
    size_t end_pos = header.substr(start_pos).find(" ");
    if (end_pos == std::string::npos) header_error();
    return (uint64_t)std::atoi(header.substr(start_pos).substr(0, end_pos).c_str());
}
inline uint64_t get_number_of_patterns(const std::string& header) { return header_field(header, "number="); }
inline uint64_t get_patterns_length(const std::string& header) { return header_field(header, "length="); }

}  // namespace ri
