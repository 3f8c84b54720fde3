// pwm.hpp -- PWM and threshold files (pattern.rs:13-117)
// Host side of find-tfbs-b200 (see driver.cpp for the map); header-only, one translation unit.
#pragma once
#include "options.hpp"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// PWMs (pattern.rs)
// ---------------------------------------------------------------------------------------------------------------
struct Pwm {
    std::vector<int32_t> w;  // len x 4
    std::string name;
    uint16_t pattern_id;
    int32_t min_score;
    uint8_t direction;
};

int32_t parse_weight(const std::string& s) {  // pattern.rs:13-16: f32, * 1000.0, round half away from zero
    char* e = nullptr;
    float x = strtof(s.c_str(), &e);
    if (e == s.c_str() || *e) die("called `Result::unwrap()` on an `Err` value: ParseFloatError (\"" + s + "\")");
    return (int32_t)roundf(x * 1000.0f);
}

std::vector<std::string> fields_ws(const std::string& l) {
    std::vector<std::string> f;
    std::istringstream is(l);
    std::string t;
    while (is >> t) f.push_back(t);
    return f;
}

bool parse_threshold_file(const std::string& path, float threshold, int32_t* out) {  // pattern.rs:18-35
    std::ifstream f(path);
    if (!f) die("Could not open file " + path);  // pattern.rs:115
    bool found = false;
    std::string line;
    while (std::getline(f, line)) {
        auto x = fields_ws(line);
        if (x.size() != 2) continue;
        int32_t w = parse_weight(x[0]);
        char* e = nullptr;
        float pv = strtof(x[1].c_str(), &e);
        if (e == x[1].c_str() || *e) die("Can't parse pvalue in file " + path);
        if (pv > threshold) { *out = w; found = true; }  // the last qualifying line wins
    }
    return found;
}

// Threshold tooling (the reference reads HOCOMOCO's .thr files, pattern.rs:18-35; synthetic PWMs need theirs written): the exact
// distribution of the integer score under uniform ACGT by dynamic programming over the columns, and the largest score s with
// P(score >= s) > pvalue -- the score of the last line of a .thr file that qualifies for `pvalue` (hits are windows with score > s).
int32_t score_threshold(const std::vector<int32_t>& w, double pvalue, double* tail_at = nullptr) {
    const size_t L = w.size() / 4;
    int64_t lo_sum = 0, span = 0;
    std::vector<int32_t> lo(L), hi(L);
    for (size_t c = 0; c < L; ++c) {
        lo[c] = *std::min_element(w.begin() + 4 * c, w.begin() + 4 * c + 4);
        hi[c] = *std::max_element(w.begin() + 4 * c, w.begin() + 4 * c + 4);
        lo_sum += lo[c];
        span += hi[c] - lo[c];
    }
    std::vector<double> dist((size_t)span + 1, 0.0), next((size_t)span + 1);
    dist[0] = 1.0;
    int64_t top = 0;
    for (size_t c = 0; c < L; ++c) {
        std::fill(next.begin(), next.end(), 0.0);
        for (int b = 0; b < 4; ++b) {
            const int64_t d = w[4 * c + b] - lo[c];
            for (int64_t i = 0; i <= top; ++i) next[(size_t)(d + i)] += 0.25 * dist[(size_t)i];
        }
        top += hi[c] - lo[c];
        dist.swap(next);
    }
    double tail = 0.0;
    int64_t k = 0;
    double tail_k = 1.0;
    bool found = false;
    for (int64_t j = span; j >= 0; --j) {  // tail = P(score - lo_sum >= j), accumulated from the top
        tail += dist[(size_t)j];
        if (!found && tail > pvalue) { k = j; tail_k = tail; found = true; }
    }
    if (tail_at) *tail_at = found ? tail_k : 1.0;
    return (int32_t)(k + lo_sum);
}

// One .thr file in HOCOMOCO's layout ("score<TAB>pvalue", ascending score): one line per requested p-value, holding the exact
// threshold for it and its exact tail probability, so that parse_threshold_file(path, p) returns score_threshold(w, p).
void write_threshold_file(const std::string& path, const std::vector<int32_t>& w, std::vector<double> pvalues) {
    std::sort(pvalues.begin(), pvalues.end(), std::greater<double>());  // larger p-value = lower score
    std::ofstream f(path);
    if (!f) die("Could not create " + path);
    int32_t last = INT32_MIN;
    for (double p : pvalues) {
        double tail = 0;
        const int32_t s = score_threshold(w, p, &tail);
        if (s == last) continue;
        last = s;
        char line[96];
        snprintf(line, sizeof line, "%.3f\t%.17g\n", s / 1000.0, tail);
        f << line;
    }
}

// Every PWM of a definition file (pattern.rs:37-87's layout), with or without a threshold: (name, weights).
std::vector<std::pair<std::string, std::vector<int32_t>>> read_pwm_definitions(const std::string& path) {
    std::ifstream f(path);
    if (!f) die("Could not open file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::vector<std::pair<std::string, std::vector<int32_t>>> out;
    for (const std::string& chunk : split(ss.str(), '>')) {
        std::vector<std::string> lines;
        for (auto& l : split(chunk, '\n'))
            if (!l.empty()) lines.push_back(l);
        if (lines.empty()) continue;
        std::vector<int32_t> w;
        for (size_t i = 1; i < lines.size(); ++i) {
            auto x = fields_ws(lines[i]);
            if (x.size() == 4)
                for (auto& t : x) w.push_back(parse_weight(t));
        }
        out.emplace_back(lines[0], w);
    }
    return out;
}

std::vector<Pwm> parse_pwm_files(const Options& o) {  // pattern.rs:37-87
    std::map<std::string, int32_t> thresholds;
    std::string dir = o.threshold_dir;
    while (!dir.empty() && dir.back() == '/') dir.pop_back();
    for (auto& p : o.pwm_names) {
        int32_t ms;
        if (parse_threshold_file(dir + "/" + p + ".thr", o.pwm_threshold, &ms)) thresholds[p] = ms;
        else printf("Could not parse %s/%s.thr\n", dir.c_str(), p.c_str());
    }
    std::ifstream f(o.pwm_file);
    if (!f) { printf("Could not open file %s\n", o.pwm_file.c_str()); exit(1); }
    std::stringstream ss;
    ss << f.rdbuf();
    std::vector<Pwm> out;
    uint16_t pattern_id = 0;
    for (const std::string& chunk : split(ss.str(), '>')) {
        if (chunk.empty()) continue;
        std::vector<std::string> lines;
        for (auto& l : split(chunk, '\n'))
            if (!l.empty()) lines.push_back(l);
        if (lines.empty()) die("index out of bounds: empty PWM definition");
        std::string name = lines[0];
        std::vector<int32_t> w;
        for (size_t i = 1; i < lines.size(); ++i) {
            auto x = fields_ws(lines[i]);
            if (x.size() == 4)
                for (auto& t : x) w.push_back(parse_weight(t));
        }
        if (std::find(o.pwm_names.begin(), o.pwm_names.end(), name) == o.pwm_names.end()) continue;
        auto it = thresholds.find(name);
        if (it == thresholds.end()) printf("Couldn't find a PWM threshold for %s\n", name.c_str());
        else {
            out.push_back(Pwm{w, name, pattern_id, it->second, TFBS_DIR_P});
            if (!o.forward_only) {  // reverse_complement, pattern.rs:103-112
                std::vector<int32_t> r(w.size());
                size_t L = w.size() / 4;
                for (size_t c = 0; c < L; ++c)
                    for (int k = 0; k < 4; ++k) r[4 * c + k] = w[4 * (L - 1 - c) + (3 - k)];
                out.push_back(Pwm{r, name, pattern_id, it->second, TFBS_DIR_N});
            }
            printf("Loaded PWM %s (len %zu, id %u, min_score %d) \n", name.c_str(), w.size() / 4, pattern_id, it->second);
        }
        pattern_id++;  // also when the threshold is missing (pattern.rs:81)
    }
    return out;
}

}  // namespace
