// test_shim.hpp -- flat entry points over the host logic, for the CPU tests (tests/test_driver_host.py).  Not part of the program.
#pragma once
#include "pwm.hpp"
#include "bed.hpp"
#include "block.hpp"
#include "rows.hpp"

extern "C" {
static thread_local std::string g_shim_err;
const char* drv_last_error() { return g_shim_err.c_str(); }

int drv_parse_weight(const char* s, int32_t* out) {
    try { *out = parse_weight(s); return 0; } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}

int drv_parse_threshold_file(const char* path, float thr, int32_t* out) {
    try { return parse_threshold_file(path, thr, out) ? 1 : 0; } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}

int drv_score_threshold(const int32_t* w, uint32_t len, double pvalue, int32_t* out) {
    try { *out = score_threshold(std::vector<int32_t>(w, w + 4 * (size_t)len), pvalue); return 0; } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}
int drv_write_threshold_file(const char* path, const int32_t* w, uint32_t len, const double* pvalues, uint32_t n) {
    try { write_threshold_file(path, std::vector<int32_t>(w, w + 4 * (size_t)len), std::vector<double>(pvalues, pvalues + n)); return 0; }
    catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}

// merge_ranges: returns the number of merged ranges written to out_s / out_e (capacity n)
int drv_merge_ranges(const uint64_t* s, const uint64_t* e, uint32_t n, uint64_t* out_s, uint64_t* out_e) {
    std::vector<Range> raw;
    for (uint32_t i = 0; i < n; ++i) raw.push_back(Range{s[i], e[i]});
    std::vector<Range> m = merge_ranges(raw);
    for (size_t i = 0; i < m.size(); ++i) { out_s[i] = m[i].start; out_e[i] = m[i].end; }
    return (int)m.size();
}

// second half of counts_as_genotypes + the maf filter: 1 = row kept (info and genotypes filled), 0 = dropped
int drv_finalise_row(const uint32_t* l, const uint32_t* r, uint32_t S, uint32_t min_maf, char* info, size_t info_cap, char* gt, size_t gt_cap) {
    uint32_t lo = UINT32_MAX, hi = 0;
    for (uint32_t i = 0; i < S; ++i) { uint32_t x = l[i] + r[i]; lo = std::min(lo, x); hi = std::max(hi, x); }
    RowText t = finalise_row(l, r, S, lo, hi, min_maf);
    if (!t.keep) return 0;
    snprintf(info, info_cap, "%s", t.info.c_str());
    snprintf(gt, gt_cap, "%s", t.genotypes.c_str());
    return 1;
}

// load_bcf: positions, allele counts and the carrier bit rows of the biallelic records of one chromosome
int drv_load_bcf(const char* bcf, const char* samples_file, const char* chrom, uint32_t cap, int64_t* pos, uint32_t* n_allele, uint32_t* carrier_row,
                 uint32_t* carriers, uint32_t carriers_cap, uint32_t* n_records, uint32_t* n_samples, uint32_t* pitch, int use_index) {
    try {
        Options o;
        o.use_index = use_index != 0;
        o.bcf = bcf;
        o.chromosome = chrom;
        if (samples_file && *samples_file) { o.has_samples = true; o.samples_file = samples_file; }
        o.threads = 4;  // exercises the parallel BGZF path
        Cohort co = load_bcf(o);
        *n_records = (uint32_t)co.records.size();
        *n_samples = (uint32_t)co.samples.size();
        *pitch = co.pitch;
        for (uint32_t i = 0; i < co.records.size() && i < cap; ++i) {
            pos[i] = co.records[i].pos;
            n_allele[i] = co.records[i].n_allele;
            carrier_row[i] = co.records[i].carrier_row;
        }
        for (size_t i = 0; i < co.carriers.size() && i < carriers_cap; ++i) carriers[i] = co.carriers[i];
        return 0;
    } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}

int drv_write_bgzf(const char* path, const char* data, uint64_t n, uint32_t threads, uint32_t piece) {
    try {
        BgzfWriter w(path, threads);
        for (uint64_t p = 0; p < n; p += piece) w.write(std::string(data + p, (size_t)std::min<uint64_t>(piece, n - p)));
        w.finish();
        return 0;
    } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}

int drv_fasta_fetch(const char* path, const char* chrom, uint64_t start, uint64_t stop, uint8_t* out, uint64_t cap, uint64_t* n) {
    try {
        Fasta fa(path, chrom);
        std::vector<uint8_t> v;
        fa.fetch(start, stop, &v);
        *n = v.size();
        memcpy(out, v.data(), std::min<uint64_t>(cap, v.size()));
        return 0;
    } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}

// parse_pwm_files: number of patterns; lens / min_scores / pattern_ids / directions (capacity cap), weights flattened (capacity wcap)
int drv_parse_pwms(const char* pwm_file, const char* thr_dir, float thr, const char* names_csv, int forward_only, uint32_t cap, uint32_t* lens,
                   int32_t* min_scores, uint16_t* pids, uint8_t* dirs, int32_t* weights, uint32_t wcap) {
    try {
        Options o;
        o.pwm_file = pwm_file;
        o.threshold_dir = thr_dir;
        o.pwm_threshold = thr;
        o.pwm_names = split(names_csv, ',');
        o.forward_only = forward_only != 0;
        std::vector<Pwm> ps = parse_pwm_files(o);
        uint32_t w = 0;
        for (uint32_t i = 0; i < ps.size() && i < cap; ++i) {
            lens[i] = (uint32_t)(ps[i].w.size() / 4);
            min_scores[i] = ps[i].min_score;
            pids[i] = ps[i].pattern_id;
            dirs[i] = ps[i].direction;
            for (int32_t x : ps[i].w)
                if (w < wcap) weights[w++] = x;
        }
        return (int)ps.size();
    } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}
}  // extern "C"
