// bed.hpp -- BED loading (bed.rs:9-60) and range merging (range.rs:43-87)
// Host side of find-tfbs-b200 (see driver.cpp for the map); header-only, one translation unit.
#pragma once
#include "options.hpp"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// BED (bed.rs) + merge (range.rs)
// ---------------------------------------------------------------------------------------------------------------
std::vector<Range> load_bed(const std::string& path, const std::string& chrom) {
    std::ifstream f(path);
    if (!f) die("Bed file " + path + " does not exist");
    std::vector<Range> xs;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        auto fld = split(line, '\t');
        if (fld.size() < 3) die("malformed BED line in " + path + ": " + line);
        char *e1 = nullptr, *e2 = nullptr;
        uint64_t s = strtoull(fld[1].c_str(), &e1, 10), e = strtoull(fld[2].c_str(), &e2, 10);
        if (fld[1].empty() || fld[2].empty() || *e1 || *e2) die("malformed BED line in " + path + ": " + line);
        if (fld[0] == chrom) xs.push_back(Range{s, e});  // start/end used as an inclusive range (bed.rs:15)
    }
    return xs;
}

std::vector<Range> merge_ranges(std::vector<Range> raw) {  // range.rs:43-87
    std::stable_sort(raw.begin(), raw.end(), [](const Range& a, const Range& b) { return a.start < b.start; });
    std::vector<Range> out;
    for (const Range& r : raw) {
        if (!out.empty() && out.back().overlaps(r)) {
            out.back().start = std::min(out.back().start, r.start);
            out.back().end = std::max(out.back().end, r.end);
        } else out.push_back(r);
    }
    return out;
}

std::string basename_of(const std::string& s) {
    size_t p = s.find_last_of('/');
    return p == std::string::npos ? s : s.substr(p + 1);
}

}  // namespace
