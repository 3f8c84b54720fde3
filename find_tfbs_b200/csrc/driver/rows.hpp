// rows.hpp -- second half of counts_as_genotypes (main.rs:459-498)
// Host side of find-tfbs-b200 (see driver.cpp for the map); header-only, one translation unit.
#pragma once
#include "common.hpp"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// rows: second half of counts_as_genotypes (main.rs:459-498) and the row text (main.rs:415-425)
// ---------------------------------------------------------------------------------------------------------------
struct RowText {
    bool keep = false;
    std::string info, genotypes;
};

template <class T>
RowText finalise_row(const T* l, const T* r, uint32_t S, uint32_t lowest, uint32_t highest, uint32_t min_maf) {
    RowText t;
    if (lowest == highest) return t;  // main.rs:456-458 (the library already filtered these)
    const uint32_t i1 = (lowest * 1000u * 3u + highest * 1000u) / 4u;  // :461
    const uint32_t i3 = (lowest * 1000u + highest * 1000u * 3u) / 4u;  // :462
    std::vector<uint32_t> all{lowest, highest};
    uint32_t zero = 0, one = 0, two = 0;
    const float lowest_f = (float)lowest, spread = (float)highest - lowest_f;
    t.genotypes.reserve((size_t)S * 12);
    char buf[48];
    for (uint32_t s = 0; s < S; ++s) {
        uint32_t x = (uint32_t)l[s] + (uint32_t)r[s];
        if (x == lowest) { t.genotypes += "\t0|0:0.0"; ++zero; }
        else if (x == highest) { t.genotypes += "\t1|1:2.0"; ++two; }
        else {
            if (std::find(all.begin(), all.end(), x) == all.end()) all.push_back(x);
            uint32_t x1000 = x * 1000u;
            if (x1000 < i1) { t.genotypes += "\t0|0"; ++zero; }
            else if (x1000 < i3) { t.genotypes += "\t0|1"; ++one; }
            else { t.genotypes += "\t1|1"; ++two; }
            volatile float num = ((float)x - lowest_f) * 2.0f;  // f32 steps as in :478
            float dosage = num / spread;
            snprintf(buf, sizeof buf, ":%.4f", (double)dosage);  // {:.4}
            t.genotypes += buf;
        }
    }
    uint32_t maf = (zero >= one && zero >= two) ? one + two : (two >= zero && two >= one) ? zero + one : zero + two;  // :482-489
    if (maf < min_maf) return t;  // main.rs:421
    std::sort(all.begin(), all.end());
    t.info = "COUNTS=";
    for (size_t i = 0; i < all.size(); ++i) t.info += (i ? "," : "") + std::to_string(all[i]);
    t.info += ";freqs=" + std::to_string(zero) + "/" + std::to_string(one) + "/" + std::to_string(two);
    t.keep = true;
    return t;
}

}  // namespace
