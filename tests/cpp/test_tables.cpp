// CPU check of the product's host-side pattern compiler (find_tfbs_b200/csrc/tables.cpp): the packed pair tables must flag a
// window iff its exact i32 score exceeds min_score (reference src/pattern.rs:125-129,151), for every pattern length 1..32, with
// N bases (score 0, src/types.rs:110), in both field formats, with several chunks, and with always/never-hit thresholds.
// The arithmetic below is what k_scan does per lane: acc = sum_g table[triple][g][pair_entry(code[i+2g], code[i+2g+1])].
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../find_tfbs_b200/csrc/tables.hpp"

using namespace tfbs;

static int check(const std::vector<tfbs_pattern>& pats, uint32_t budget, int force_wide, std::mt19937_64& rng, const char* what) {
    CompiledPatterns cp;
    std::string err;
    int rc = compile_patterns(pats.data(), (uint32_t)pats.size(), budget, force_wide, &cp, &err);
    if (rc != TFBS_OK) { printf("FAIL %s: compile error %d %s\n", what, rc, err.c_str()); return 1; }
    const int bits = cp.fields == 3 ? 21 : 32;
    std::vector<uint8_t> seq(600);
    for (auto& c : seq) c = (rng() % 23 == 0) ? 4 : (uint8_t)(rng() % 4);
    std::vector<int> seen(pats.size(), 0);
    long long n_hits = 0;
    for (const ChunkDesc& cd : cp.chunks) {
        size_t word = cd.tbl_off;
        uint32_t t = 0;
        for (uint32_t rn = 0; rn < cd.n_runs; ++rn) {
            const RunDesc rd = cp.runs[cd.run_off + rn];
            for (uint32_t k = 0; k < rd.n_triples; ++k, ++t) {
                for (size_t i = 0; i < seq.size(); ++i) {
                    uint64_t acc = 0;
                    for (uint32_t g = 0; g < rd.groups; ++g) {
                        int a = i + 2 * g < seq.size() ? seq[i + 2 * g] : 4, b = i + 2 * g + 1 < seq.size() ? seq[i + 2 * g + 1] : 4;
                        acc += cp.table[word + (size_t)g * kPairEntries + pair_entry(a, b)];
                    }
                    for (uint32_t f = 0; f < cp.fields; ++f) {
                        int pi = cp.trip_pat[(size_t)(cd.trip_off + t) * 3 + f];
                        bool flag = (acc >> (bits * f + bits - 1)) & 1;
                        if (pi < 0) { if (flag) { printf("FAIL %s: empty slot flagged\n", what); return 1; } continue; }
                        const tfbs_pattern& p = pats[pi];
                        if (i == 0) seen[pi]++;
                        if (i + p.len > seq.size()) continue;  // incomplete windows are filtered in the rare path
                        long long score = 0;
                        for (uint32_t c = 0; c < p.len; ++c) score += seq[i + c] < 4 ? p.weights[4 * c + seq[i + c]] : 0;
                        bool hit = score > p.min_score;
                        n_hits += hit;
                        if (hit != flag) {
                            printf("FAIL %s: pattern %d len %u window %zu score %lld min_score %d flag %d\n", what, pi, p.len, i, score, p.min_score, (int)flag);
                            return 1;
                        }
                    }
                }
                word += (size_t)rd.groups * kPairEntries;
            }
        }
        if (t != cd.n_triples) { printf("FAIL %s: run/triple bookkeeping\n", what); return 1; }
        if (cd.tbl_off % 2) { printf("FAIL %s: chunk not 16-byte aligned\n", what); return 1; }
    }
    for (size_t i = 0; i < pats.size(); ++i)
        if (pats[i].kind == TFBS_PATTERN_PWM && seen[i] != 1) { printf("FAIL %s: pattern %zu appears %d times\n", what, i, seen[i]); return 1; }
    printf("ok %s: %zu patterns, %zu chunks, fields %u, %lld hits\n", what, pats.size(), cp.chunks.size(), cp.fields, n_hits);
    return 0;
}

int main() {
    std::mt19937_64 rng(12345);
    int bad = 0;
    for (int round = 0; round < 6; ++round) {
        std::vector<std::vector<int32_t>> store;
        std::vector<tfbs_pattern> pats;
        int n = 5 + (int)(rng() % 60);
        int scale = round == 3 ? 60000 : (round == 4 ? 3 : 6000);  // 3: too wide for 21-bit fields; 4: tiny weights
        for (int k = 0; k < n; ++k) {
            uint32_t L = round == 5 ? 1 + (uint32_t)(k % 32) : 1 + (uint32_t)(rng() % 32);
            std::vector<int32_t> w(4 * L);
            long long best = 0, worst = 0;
            for (uint32_t c = 0; c < L; ++c) {
                int mx = -1000000000, mn = 1000000000;
                for (int x = 0; x < 4; ++x) {
                    int v = (int)((long long)(rng() % (2 * scale + 1)) - scale - scale / 3);
                    w[4 * c + x] = v;
                    mx = v > mx ? v : mx;
                    mn = v < mn ? v : mn;
                }
                best += mx > 0 ? mx : 0;  // N scores 0
                worst += mn < 0 ? mn : 0;
            }
            store.push_back(w);
            tfbs_pattern p{};
            p.weights = store.back().data();
            p.len = L;
            long long span = best - worst;
            int mode = (int)(rng() % 10);
            p.min_score = mode == 0 ? (int)(worst - 5) : mode == 1 ? (int)(best + 5) : mode == 2 ? (int)best - 1 : (int)(worst + span * (30 + (long long)(rng() % 60)) / 100);
            p.pattern_id = (uint16_t)(k / 2);
            p.direction = (uint8_t)(k & 1);
            p.kind = TFBS_PATTERN_PWM;
            pats.push_back(p);
        }
        tfbs_pattern other{};
        other.kind = TFBS_PATTERN_OTHER;
        other.pattern_id = 999;
        pats.push_back(other);
        for (size_t i = 0; i + 1 < pats.size(); ++i) pats[i].weights = store[i].data();
        char name[64];
        snprintf(name, sizeof name, "round %d (auto)", round);
        bad += check(pats, 96 * 1024, 0, rng, name);
        snprintf(name, sizeof name, "round %d (wide, 12 KB chunks)", round);
        bad += check(pats, 12 * 1024, 1, rng, name);
    }
    // error paths
    {
        CompiledPatterns cp;
        std::string err;
        std::vector<int32_t> w(4 * 33, 1);
        tfbs_pattern p{w.data(), 33, 0, 0, 0, TFBS_PATTERN_PWM};
        if (compile_patterns(&p, 1, 96 * 1024, 0, &cp, &err) != TFBS_ERR_INVALID_ARGUMENT) { printf("FAIL: 33 columns accepted\n"); ++bad; }
        tfbs_pattern e{nullptr, 0, 0, 0, 0, TFBS_PATTERN_PWM};
        if (compile_patterns(&e, 1, 96 * 1024, 0, &cp, &err) != TFBS_ERR_INVALID_ARGUMENT) { printf("FAIL: empty PWM accepted\n"); ++bad; }
        std::vector<int32_t> big(4 * 20);
        for (size_t i = 0; i < big.size(); ++i) big[i] = (i % 4 == 0) ? 2000000000 : -2000000000;
        tfbs_pattern g{big.data(), 20, 0, 0, 0, TFBS_PATTERN_PWM};
        if (compile_patterns(&g, 1, 96 * 1024, 0, &cp, &err) != TFBS_ERR_SCORE_RANGE) { printf("FAIL: score range not detected\n"); ++bad; }
    }
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
