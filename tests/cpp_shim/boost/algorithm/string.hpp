// Test-only stand-in for <boost/algorithm/string.hpp> (not installed here): split + is_any_of, the two names the reference's
// CSVReader.h uses, so that main.cpp / experiment.cpp compile against the drop-in headers.
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of_t { std::string set; };
inline is_any_of_t is_any_of(const std::string& s) { return is_any_of_t{s}; }
namespace algorithm {
inline void split(std::vector<std::string>& out, const std::string& line, const is_any_of_t& sep) {
    out.clear();
    std::string cur;
    for (char c : line) { if (sep.set.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); } else cur.push_back(c); }
    out.push_back(cur);
}
}  // namespace algorithm
}  // namespace boost
