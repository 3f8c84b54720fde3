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
