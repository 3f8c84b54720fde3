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

// Text of one sample for a count x (main.rs:463-481): the class by the integer thresholds i1 / i3, the dosage with f32 steps and {:.4}.
struct SampleText {
    char txt[16];  // "\t0|1:0.6667" is the longest (11 characters)
    uint8_t len = 0, cls = 0;
};
inline SampleText sample_text(uint32_t x, uint32_t lowest, uint32_t highest, uint32_t i1, uint32_t i3) {
    SampleText e;
    if (x == lowest) { memcpy(e.txt, "\t0|0:0.0", 8); e.len = 8; e.cls = 0; return e; }
    if (x == highest) { memcpy(e.txt, "\t1|1:2.0", 8); e.len = 8; e.cls = 2; return e; }
    const uint32_t x1000 = x * 1000u;
    e.cls = x1000 < i1 ? 0 : (x1000 < i3 ? 1 : 2);
    const float lowest_f = (float)lowest, spread = (float)highest - lowest_f;
    volatile float num = ((float)x - lowest_f) * 2.0f;  // f32 steps as in :478
    const float dosage = num / spread;
    e.len = (uint8_t)snprintf(e.txt, sizeof e.txt, "\t%s:%.4f", e.cls == 0 ? "0|0" : (e.cls == 1 ? "0|1" : "1|1"), (double)dosage);  // {:.4}
    return e;
}

template <class T>
RowText finalise_row(const T* l, const T* r, uint32_t S, uint32_t lowest, uint32_t highest, uint32_t min_maf) {
    RowText t;
    if (lowest == highest) return t;  // main.rs:456-458 (the library already filtered these)
    const uint32_t i1 = (lowest * 1000u * 3u + highest * 1000u) / 4u;  // :461
    const uint32_t i3 = (lowest * 1000u + highest * 1000u * 3u) / 4u;  // :462
    uint32_t cls_n[3] = {0, 0, 0};
    std::vector<uint32_t> all;
    const uint32_t span = highest - lowest;
    t.genotypes.resize((size_t)S * 12 + 16);
    char* const g0 = &t.genotypes[0];
    char* o = g0;
    if (span < 1024) {
        // a row holds a handful of distinct counts: the text of each is made once, a sample costs one table entry and a 16-byte copy
        SampleText table[1024];
        for (uint32_t s = 0; s < S; ++s) {
            const uint32_t x = (uint32_t)l[s] + (uint32_t)r[s];
            if (x - lowest > span) die("internal error: a count outside [min, max] of its row");
            SampleText& e = table[x - lowest];
            if (!e.len) e = sample_text(x, lowest, highest, i1, i3);
            memcpy(o, e.txt, 16);
            o += e.len;
            ++cls_n[e.cls];
        }
        for (uint32_t d = 0; d <= span; ++d)
            if (table[d].len || d == 0 || d == span) all.push_back(lowest + d);
    } else {
        all = {lowest, highest};
        for (uint32_t s = 0; s < S; ++s) {
            const uint32_t x = (uint32_t)l[s] + (uint32_t)r[s];
            const SampleText e = sample_text(x, lowest, highest, i1, i3);
            if (std::find(all.begin(), all.end(), x) == all.end()) all.push_back(x);
            memcpy(o, e.txt, 16);
            o += e.len;
            ++cls_n[e.cls];
        }
        std::sort(all.begin(), all.end());
    }
    t.genotypes.resize((size_t)(o - g0));
    const uint32_t zero = cls_n[0], one = cls_n[1], two = cls_n[2];
    uint32_t maf = (zero >= one && zero >= two) ? one + two : (two >= zero && two >= one) ? zero + one : zero + two;  // :482-489
    if (maf < min_maf) return t;  // main.rs:421
    t.info = "COUNTS=";
    for (size_t i = 0; i < all.size(); ++i) t.info += (i ? "," : "") + std::to_string(all[i]);
    t.info += ";freqs=" + std::to_string(zero) + "/" + std::to_string(one) + "/" + std::to_string(two);
    t.keep = true;
    return t;
}

}  // namespace
