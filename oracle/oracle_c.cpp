// oracle_c.cpp -- flat C entry points over the CPU oracle so tests/ and bench.py can drive it with
// ctypes.  TEST INFRASTRUCTURE ONLY (see oracle.hpp): the product never links this.
#include <cstdlib>
#include <cstring>
#include <string>

#include "oracle.hpp"

using namespace ora;

namespace {
thread_local std::string g_err;
int fail(const OracleError& e) {
    g_err = e.message;
    return e.code ? e.code : TFBS_ERR_INTERNAL;
}
template <class T>
T* dup(const std::vector<T>& v) {
    T* p = (T*)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}
char* dup_str(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
std::vector<std::string> split(const char* s, char sep) {
    std::vector<std::string> out;
    std::string cur;
    for (const char* p = s; *p; ++p) {
        if (*p == sep) { out.push_back(cur); cur.clear(); }
        else cur += *p;
    }
    out.push_back(cur);
    return out;
}
std::vector<Nucleotide> nucs(const char* s) { return to_nucleotides((const uint8_t*)s, strlen(s)); }
}  // namespace

extern "C" {

const char* ora_last_error(void) { return g_err.c_str(); }
void ora_free(void* p) { free(p); }
void ora_set_chunk_size(uint32_t n) { set_chunk_size(n); }

// ---- block level ------------------------------------------------------------------------------
typedef struct ora_block_result {
    uint64_t n_rows;
    uint32_t* region;
    uint32_t* inner;
    uint16_t* pattern_id;
    uint32_t* vmin;
    uint32_t* vmax;
    uint32_t* left;   // n_rows * n_samples
    uint32_t* right;
    uint64_t n_matches;
    uint32_t* m_region;
    uint32_t* m_pattern_index;
    uint32_t* m_group;
    int64_t* m_start;
    uint64_t n_hap_group;
    uint32_t* hap_group;
    uint64_t executed_cells, nominal_cells, n_groups, n_hits;
    uint32_t collision_regions, truncated_regions;
    uint64_t n_hap_flags;
    uint8_t* hap_flags;
} ora_block_result;

int ora_process_block(const tfbs_pattern* patterns, uint32_t n_patterns, const tfbs_block* blk, int rows_mode, int want_matches,
                      int n_threads, ora_block_result* out) {
    memset(out, 0, sizeof *out);
    try {
        std::vector<Pattern> pl = patterns_from_c(patterns, n_patterns);
        BlockResult r;
        process_block(pl, *blk, rows_mode, want_matches != 0, n_threads, &r);
        const uint32_t S = blk->n_samples;
        out->n_rows = r.rows.size();
        std::vector<uint32_t> region, inner, vmin, vmax, left, right;
        std::vector<uint16_t> pid;
        for (const BlockRow& b : r.rows) {
            region.push_back(b.region);
            inner.push_back(b.inner);
            pid.push_back(b.pattern_id);
            vmin.push_back(b.vmin);
            vmax.push_back(b.vmax);
            left.insert(left.end(), b.left.begin(), b.left.end());
            right.insert(right.end(), b.right.begin(), b.right.end());
        }
        (void)S;
        out->region = dup(region);
        out->inner = dup(inner);
        out->pattern_id = dup(pid);
        out->vmin = dup(vmin);
        out->vmax = dup(vmax);
        out->left = dup(left);
        out->right = dup(right);
        std::vector<uint32_t> mr, mp, mg;
        std::vector<int64_t> ms;
        for (const BlockMatch& m : r.matches) {
            mr.push_back(m.region);
            mp.push_back(m.pattern_index);
            mg.push_back(m.group);
            ms.push_back(m.start);
        }
        out->n_matches = r.matches.size();
        out->m_region = dup(mr);
        out->m_pattern_index = dup(mp);
        out->m_group = dup(mg);
        out->m_start = dup(ms);
        out->n_hap_group = r.hap_group.size();
        out->hap_group = dup(r.hap_group);
        out->executed_cells = r.executed_cells;
        out->nominal_cells = r.nominal_cells;
        out->n_groups = r.n_groups;
        out->n_hits = r.n_hits;
        out->collision_regions = r.collision_regions;
        out->truncated_regions = r.truncated_regions;
        out->n_hap_flags = r.hap_flags.size();
        out->hap_flags = dup(r.hap_flags);
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

void ora_free_block_result(ora_block_result* r) {
    free(r->region); free(r->inner); free(r->pattern_id); free(r->vmin); free(r->vmax); free(r->left); free(r->right);
    free(r->m_region); free(r->m_pattern_index); free(r->m_group); free(r->m_start); free(r->hap_group); free(r->hap_flags);
    memset(r, 0, sizeof *r);
}

// ---- unit level (the reference's own test vectors) ----------------------------------------------

// patch_haplotype: window bases `ref` (ASCII) with explicit positions; diffs as parallel arrays of
// pos and NUL-terminated allele strings.  Output: codes (0..4) and positions, at most cap entries.
int ora_patch_haplotype(uint64_t range_start, uint64_t range_end, uint32_t n_diffs, const uint64_t* diff_pos,
                        const char* const* diff_ref, const char* const* diff_alt, uint32_t n_ref, const char* ref_letters,
                        const uint64_t* ref_pos, uint32_t cap, uint8_t* out_nuc, uint64_t* out_pos, uint32_t* out_len,
                        int* truncated) {
    try {
        std::vector<Diff> diffs;
        for (uint32_t i = 0; i < n_diffs; ++i) diffs.push_back(Diff{diff_pos[i], nucs(diff_ref[i]), nucs(diff_alt[i])});
        std::vector<NucleotidePos> ref;
        for (uint32_t i = 0; i < n_ref; ++i) ref.push_back(NucleotidePos{to_nucleotide((uint8_t)ref_letters[i]), ref_pos[i]});
        bool tr = false;
        std::vector<NucleotidePos> p = patch_haplotype(Range{range_start, range_end}, diffs, ref, &tr);
        *out_len = (uint32_t)p.size();
        if (truncated) *truncated = tr;
        for (uint32_t i = 0; i < p.size() && i < cap; ++i) {
            out_nuc[i] = p[i].nuc;
            out_pos[i] = p[i].pos;
        }
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

// matches(): one pattern (weights len x 4) over a haplotype given as letters + positions.
int ora_matches(const int32_t* weights, uint32_t len, int32_t min_score, uint16_t pattern_id, uint32_t n, const char* letters,
                const uint64_t* pos, uint32_t cap, uint64_t* out_start, uint64_t* out_end, uint16_t* out_pid, uint32_t* out_n) {
    try {
        Pattern p;
        p.pattern_id = pattern_id;
        p.min_score = min_score;
        for (uint32_t c = 0; c < len; ++c) p.weights.push_back(Weight::make(weights[4 * c], weights[4 * c + 1], weights[4 * c + 2], weights[4 * c + 3]));
        std::vector<NucleotidePos> h;
        for (uint32_t i = 0; i < n; ++i) h.push_back(NucleotidePos{to_nucleotide((uint8_t)letters[i]), pos[i]});
        std::vector<Match> ms;
        matches(p, 0, h, 0, &ms);
        *out_n = (uint32_t)ms.size();
        for (uint32_t i = 0; i < ms.size() && i < cap; ++i) {
            out_start[i] = ms[i].range.start;
            out_end[i] = ms[i].range.end;
            out_pid[i] = ms[i].pattern_id;
        }
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

int ora_parse_weight(const char* s, int32_t* out) {
    try {
        *out = parse_weight(s);
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

void ora_reverse_complement(const int32_t* w, uint32_t len, int32_t* out) {
    std::vector<Weight> ws;
    for (uint32_t c = 0; c < len; ++c) ws.push_back(Weight::make(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
    std::vector<Weight> r = reverse_complement(ws);
    for (uint32_t c = 0; c < len; ++c)
        for (int k = 0; k < 4; ++k) out[4 * c + k] = r[c].acgtn[k];
}

int ora_parse_threshold_file(const char* path, float threshold, int32_t* out, int* found) {
    try {
        *found = parse_threshold_file(path, threshold, out) ? 1 : 0;
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

// parse_pwm_files: returns flat arrays (malloc'd; free with ora_free).  names are '\n'-joined.
int ora_parse_pwm_files(const char* pwm_file, const char* threshold_dir, float threshold, const char* wanted_csv, int add_reverse,
                        uint32_t* n_out, uint32_t** len_out, int32_t** weights_out, uint16_t** pid_out, int32_t** min_score_out,
                        uint8_t** dir_out, char** names_out) {
    try {
        std::vector<Pattern> ps = parse_pwm_files(pwm_file, threshold_dir, threshold, split(wanted_csv, ','), add_reverse != 0);
        std::vector<uint32_t> len;
        std::vector<int32_t> w, ms;
        std::vector<uint16_t> pid;
        std::vector<uint8_t> dir;
        std::string names;
        for (const Pattern& p : ps) {
            len.push_back((uint32_t)p.weights.size());
            for (const Weight& c : p.weights)
                for (int k = 0; k < 4; ++k) w.push_back(c.acgtn[k]);
            ms.push_back(p.min_score);
            pid.push_back(p.pattern_id);
            dir.push_back(p.direction);
            names += p.name + "\n";
        }
        *n_out = (uint32_t)ps.size();
        *len_out = dup(len);
        *weights_out = dup(w);
        *pid_out = dup(pid);
        *min_score_out = dup(ms);
        *dir_out = dup(dir);
        *names_out = dup_str(names);
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

// count_matches_by_sample on explicit matches with single-haplotype carrier lists
// (main.rs:570-671 vectors).  Output rows: (bed, start, end, pattern_id, left[S], right[S]).
int ora_count_matches(uint32_t n_matches, const uint64_t* m_start, const uint64_t* m_end, const uint16_t* m_pid,
                      const uint32_t* m_hap /* 2*sample+side */, uint32_t n_inner, const uint32_t* in_bed, const uint64_t* in_start,
                      const uint64_t* in_end, uint32_t sample_count, uint32_t cap, uint32_t* out_bed, uint64_t* out_start,
                      uint64_t* out_end, uint16_t* out_pid, uint32_t* out_left, uint32_t* out_right, uint32_t* out_n) {
    RegionMatches rm;
    rm.group_ids.emplace_back();
    for (uint32_t i = 0; i < n_matches; ++i) {
        rm.group_ids.push_back({m_hap[i]});
        rm.match_list.push_back(Match{Range{m_start[i], m_end[i]}, m_pid[i], 0, (uint32_t)rm.group_ids.size() - 1});
    }
    std::vector<InnerPeak> inner;
    for (uint32_t k = 0; k < n_inner; ++k) inner.push_back(InnerPeak{in_bed[k], Range{in_start[k], in_end[k]}, k});
    std::map<CountKey, CountValue> c = count_matches_by_sample(rm, inner, sample_count);
    *out_n = (uint32_t)c.size();
    uint32_t i = 0;
    for (auto& kv : c) {
        if (i >= cap) break;
        out_bed[i] = kv.first.bed_index;
        out_start[i] = kv.first.range.start;
        out_end[i] = kv.first.range.end;
        out_pid[i] = kv.first.pattern_id;
        memcpy(out_left + (size_t)i * sample_count, kv.second.left.data(), sample_count * 4);
        memcpy(out_right + (size_t)i * sample_count, kv.second.right.data(), sample_count * 4);
        ++i;
    }
    return 0;
}

// counts_as_genotypes: returns 0 if the row is dropped (min == max), 1 otherwise.
int ora_counts_as_genotypes(const uint32_t* v1, const uint32_t* v2, uint32_t n, uint32_t* maf, uint32_t* freqs /*3*/,
                            char** counts_csv, char** genotypes) {
    GenotypeRow g;
    std::vector<uint32_t> a(v1, v1 + n), b(v2, v2 + n);
    if (!counts_as_genotypes(a, b, &g)) return 0;
    *maf = g.maf;
    freqs[0] = g.freq0;
    freqs[1] = g.freq1;
    freqs[2] = g.freq2;
    std::string s;
    for (size_t i = 0; i < g.distinct_counts.size(); ++i) s += (i ? "," : "") + std::to_string(g.distinct_counts[i]);
    *counts_csv = dup_str(s);
    *genotypes = dup_str(g.genotypes);
    return 1;
}

// load_peak_files: merged regions + per-file kept regions (flattened with offsets).
int ora_load_peak_files(const char* beds_csv, const char* chromosome, uint64_t after_position, uint32_t* n_merged,
                        uint64_t** merged_start, uint64_t** merged_end, uint32_t* n_files, uint32_t** file_off,
                        uint64_t** file_start, uint64_t** file_end, char** names) {
    try {
        std::vector<Range> merged;
        std::vector<std::vector<Range>> pm;
        std::vector<std::string> bn;
        load_peak_files(split(beds_csv, ','), chromosome, after_position, &merged, &pm, &bn);
        std::vector<uint64_t> ms, me, fs, fe;
        std::vector<uint32_t> off{0};
        for (const Range& r : merged) { ms.push_back(r.start); me.push_back(r.end); }
        std::string nm;
        for (size_t b = 0; b < pm.size(); ++b) {
            for (const Range& r : pm[b]) { fs.push_back(r.start); fe.push_back(r.end); }
            off.push_back((uint32_t)fs.size());
            nm += bn[b] + "\n";
        }
        *n_merged = (uint32_t)merged.size();
        *merged_start = dup(ms);
        *merged_end = dup(me);
        *n_files = (uint32_t)pm.size();
        *file_off = dup(off);
        *file_start = dup(fs);
        *file_end = dup(fe);
        *names = dup_str(nm);
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

int ora_range_overlaps(uint64_t s0, uint64_t e0, uint64_t s1, uint64_t e1) { return Range{s0, e0}.overlaps(Range{s1, e1}) ? 1 : 0; }
int ora_range_contains(uint64_t s0, uint64_t e0, uint64_t p) { return Range{s0, e0}.contains(p) ? 1 : 0; }

// select_inner_peaks over a flattened peak_map; returns indices into the flattened arrays.
int ora_select_inner_peaks(uint64_t m_start, uint64_t m_end, uint32_t n_files, const uint32_t* file_off, const uint64_t* fs,
                           const uint64_t* fe, uint32_t cap, uint32_t* out_idx, uint32_t* out_n) {
    std::vector<std::vector<Range>> pm(n_files);
    for (uint32_t b = 0; b < n_files; ++b)
        for (uint32_t k = file_off[b]; k < file_off[b + 1]; ++k) pm[b].push_back(Range{fs[k], fe[k]});
    std::vector<InnerPeak> ip = select_inner_peaks(Range{m_start, m_end}, pm);
    *out_n = (uint32_t)ip.size();
    uint32_t i = 0;
    for (const InnerPeak& p : ip) {
        if (i >= cap) break;
        // recover the flattened index of this occurrence
        uint32_t seen = 0;
        for (uint32_t k = file_off[p.bed_index]; k < file_off[p.bed_index + 1]; ++k)
            if (fs[k] == p.range.start && fe[k] == p.range.end) {
                bool used = false;
                for (uint32_t j = 0; j < i; ++j) used |= out_idx[j] == k;
                if (!used) { out_idx[i] = k; seen = 1; break; }
            }
        if (!seen) out_idx[i] = UINT32_MAX;
        ++i;
    }
    return 0;
}

// run(): the whole program on files; returns the decompressed VCF text (malloc'd).
int ora_run(const char* chromosome, const char* bcf, const char* beds_csv, const char* reference, const char* samples_file,
            const char* pwm_file, const char* threshold_dir, float threshold, const char* wanted_csv, const char* output,
            int forward_only, uint32_t min_maf, uint32_t threads, uint64_t after_position, char** text_out) {
    try {
        RunOptions o;
        o.chromosome = chromosome;
        o.bcf = bcf;
        o.bed_files = split(beds_csv, ',');
        o.reference = reference;
        o.has_samples = samples_file && *samples_file;
        if (o.has_samples) o.samples_file = samples_file;
        o.pwm_file = pwm_file;
        o.pwm_threshold_dir = threshold_dir;
        o.pwm_threshold = threshold;
        o.wanted_pwms = split(wanted_csv, ',');
        o.output = output ? output : "";
        o.forward_only = forward_only != 0;
        o.min_maf = min_maf;
        o.threads = threads;
        o.after_position = after_position;
        std::string t = run(o);
        *text_out = dup_str(t);
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

int ora_gunzip_file(const char* path, char** text_out) {
    try {
        *text_out = dup_str(gunzip_file(path));
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

// Decoded view of a BCF for fixture checks: positions, allele strings ('\n'-joined "REF,ALT,..."), raw GT codes.
int ora_read_bcf(const char* path, uint32_t* n_records, int64_t** pos, char** alleles, uint32_t* n_samples, int32_t** gt /* n_records*n_samples*2 */,
                 char** sample_names, char** contigs) {
    try {
        BcfFile bf = read_bcf(path);
        std::vector<int64_t> p;
        std::vector<int32_t> g;
        std::string al, sn, cn;
        for (const BcfRecord& r : bf.records) {
            p.push_back(r.pos);
            for (size_t a = 0; a < r.alleles.size(); ++a) al += (a ? "," : "") + r.alleles[a];
            al += "\n";
            for (size_t s = 0; s < bf.samples.size(); ++s)
                for (int k = 0; k < 2; ++k) g.push_back(r.gt_ploidy == 2 ? r.gt[s * 2 + k] : -1);
        }
        for (auto& s : bf.samples) sn += s + "\n";
        for (auto& s : bf.contigs) cn += s + "\n";
        *n_records = (uint32_t)bf.records.size();
        *pos = dup(p);
        *alleles = dup_str(al);
        *n_samples = (uint32_t)bf.samples.size();
        *gt = dup(g);
        *sample_names = dup_str(sn);
        *contigs = dup_str(cn);
        return 0;
    } catch (const OracleError& e) {
        return fail(e);
    }
}

}  // extern "C"
