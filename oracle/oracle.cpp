// oracle.cpp -- CPU restatement of the find-tfbs hot path.  TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Every function cites the reference file:line it follows.
#include "oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <set>
#include <sstream>
#include <thread>

namespace ora {

// ---------------------------------------------------------------------------------------------
// util.rs
// ---------------------------------------------------------------------------------------------

// util.rs:4-16: upper and lower case ACGTN, anything else panics.
Nucleotide to_nucleotide(uint8_t l) {
    switch (l) {
        case 65: case 97: return A;
        case 67: case 99: return C;
        case 71: case 103: return G;
        case 84: case 116: return T;
        case 78: case 110: return N;
        default:
            throw OracleError{TFBS_ERR_UNKNOWN_NUCLEOTIDE, "Unknown nucleotide " + std::to_string((int)l)};
    }
}

std::vector<Nucleotide> to_nucleotides(const uint8_t* letters, size_t n) {
    std::vector<Nucleotide> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = to_nucleotide(letters[i]);
    return v;
}

// util.rs:22-31: base i gets pos = r.start + i.
std::vector<NucleotidePos> to_nucleotides_pos(const uint8_t* letters, size_t n, const Range& r) {
    std::vector<NucleotidePos> res;
    res.reserve(n);
    uint64_t pos = r.start;
    for (size_t i = 0; i < n; ++i) res.push_back(NucleotidePos{to_nucleotide(letters[i]), pos++});
    return res;
}

// ---------------------------------------------------------------------------------------------
// haplotype.rs
// ---------------------------------------------------------------------------------------------

// haplotype.rs:90-92: every base of the window whose pos lies in [r.start, r.end].
static void get(const Range& r, const std::vector<NucleotidePos>& ref_genome_peak, std::vector<NucleotidePos>* out) {
    for (const NucleotidePos& n : ref_genome_peak)
        if (n.pos >= r.start && n.pos <= r.end) out->push_back(n);
}

static std::string diff_to_string(const Diff& d) {
    static const char L[] = "ACGTN";
    std::string s = std::to_string(d.pos) + " ";
    for (Nucleotide n : d.reference) s += L[n];
    s += "->";
    for (Nucleotide n : d.alternative) s += L[n];
    return s;
}

// haplotype.rs:94-156.  The reference recurses chunk by chunk; this loop visits the same
// cases in the same order.
std::vector<NucleotidePos> patch_haplotype(const Range& range, const std::vector<Diff>& diffs,
                                           const std::vector<NucleotidePos>& ref_haplotype, bool* truncated) {
    if (truncated) *truncated = false;
    std::vector<const Diff*> ds;  // :95-96 keep diffs starting inside the window, sorted by derived Ord
    for (const Diff& d : diffs)
        if (d.pos >= range.start && d.pos <= range.end) ds.push_back(&d);
    std::stable_sort(ds.begin(), ds.end(), [](const Diff* a, const Diff* b) { return *a < *b; });

    std::vector<NucleotidePos> out;
    // get() on a window that util.rs built is a contiguous slice; use that when it holds.
    bool contiguous = true;
    for (size_t i = 0; i < ref_haplotype.size() && contiguous; ++i)
        contiguous = ref_haplotype[i].pos == ref_haplotype[0].pos + i;
    auto get_fast = [&](uint64_t s, uint64_t e) {
        if (e < s) return;
        if (!contiguous) { get(Range{s, e}, ref_haplotype, &out); return; }
        if (ref_haplotype.empty()) return;
        uint64_t p0 = ref_haplotype[0].pos, p1 = p0 + ref_haplotype.size() - 1;
        if (e < p0 || s > p1) return;
        uint64_t a = std::max(s, p0), b = std::min(e, p1);
        out.insert(out.end(), ref_haplotype.begin() + (a - p0), ref_haplotype.begin() + (b - p0) + 1);
    };

    uint64_t ref_position = range.start;
    size_t k = 0;
    for (;;) {
        if (k == ds.size()) {  // :100-108
            if (ref_position <= range.end) get_fast(ref_position, range.end);
            return out;
        }
        const Diff& d = *ds[k];
        if (d.pos > ref_position) {  // :110-114
            get_fast(ref_position, d.pos - 1);
            ref_position = d.pos;
        } else if (d.pos == ref_position && d.reference.size() == 1) {  // :115-135 SNV or insertion
            Nucleotide at = N;  // :119-125 N if the window does not hold this position
            for (const NucleotidePos& np : ref_haplotype)
                if (np.pos == ref_position) at = np.nuc;
            if (d.reference[0] != at)
                throw OracleError{TFBS_ERR_REF_MISMATCH,
                                  std::string("First reference nucleotide of variant doesn't match reference genome: reference_first_nuc_at=") +
                                      "ACGTN"[at] + " ref_position=" + std::to_string(ref_position) + " diff=" + diff_to_string(d)};
            for (Nucleotide n : d.alternative) out.push_back(NucleotidePos{n, ref_position});  // :130-132 all share the anchor pos
            ref_position += 1;
            ++k;
        } else if (d.pos == ref_position && d.alternative.size() == 1) {  // :136-140 deletion
            out.push_back(NucleotidePos{d.alternative[0], ref_position});
            ref_position += d.reference.size();
            ++k;
        } else if (d.pos == ref_position) {  // :141-143
            throw OracleError{TFBS_ERR_MISSING_CASE, "Missing case in haplotype patcher"};
        } else if (ref_position >= range.end) {  // :144-146 overlapped variant: truncate
            if (truncated) *truncated = true;
            get_fast(ref_position, ref_position);
            return out;
        } else {  // :147-149
            if (truncated) *truncated = true;
            return out;
        }
    }
}

// haplotype.rs:65-88.  group_by_diffs inverts haplotype -> Vec<Diff>; load_haplotypes patches each
// distinct list and re-keys the result BY PATCHED SEQUENCE (:84), so two lists that patch to the
// same (nuc,pos) vector overwrite each other and the loser's haplotypes silently stay in the
// reference set (main.rs:103-105,129-131).  The reference's winner depends on HashMap iteration
// order (RandomState); the oracle fixes it: the group holding the smallest haplotype index wins.
LoadedHaplotypes load_haplotypes(const Range& peak, const std::vector<std::vector<Diff>>& diffs_by_haplotype,
                                 uint32_t variant_count, const std::vector<NucleotidePos>& ref_haplotype) {
    LoadedHaplotypes res;
    res.variant_count = variant_count;
    std::map<std::vector<Diff>, std::vector<uint32_t>> by_diffs;  // group_by_diffs :65-75
    for (uint32_t h = 0; h < diffs_by_haplotype.size(); ++h)
        if (!diffs_by_haplotype[h].empty()) by_diffs[diffs_by_haplotype[h]].push_back(h);  // only carriers have an entry (:42-49)

    struct Raw {
        const std::vector<Diff>* diffs;
        const std::vector<uint32_t>* ids;
    };
    std::vector<Raw> raw;
    for (auto& kv : by_diffs) raw.push_back(Raw{&kv.first, &kv.second});
    // ids are pushed in ascending h, so ids->front() is the smallest member.
    std::sort(raw.begin(), raw.end(), [](const Raw& a, const Raw& b) { return a.ids->front() < b.ids->front(); });

    std::map<std::vector<NucleotidePos>, size_t> by_sequence;  // :81-85
    for (const Raw& r : raw) {
        HaplotypeGroup g;
        g.sequence = patch_haplotype(peak, *r.diffs, ref_haplotype, &g.truncated);
        if (g.truncated) res.truncated_ids.insert(res.truncated_ids.end(), r.ids->begin(), r.ids->end());
        auto it = by_sequence.find(g.sequence);
        if (it != by_sequence.end()) {
            // a later insert would overwrite the earlier one in the reference; here the group with
            // the smallest first id (already stored) stays and this one is dropped.
            res.overwritten_ids.insert(res.overwritten_ids.end(), r.ids->begin(), r.ids->end());
            if (!(g.sequence == ref_haplotype)) res.sequence_collision = true;
            continue;
        }
        g.haplotype_ids = *r.ids;
        g.diffs = *r.diffs;
        by_sequence.emplace(g.sequence, res.groups.size());
        res.groups.push_back(std::move(g));
    }
    return res;
}

// ---------------------------------------------------------------------------------------------
// pattern.rs
// ---------------------------------------------------------------------------------------------

// pattern.rs:13-16: f32 parse, * 1000.0f32, round half away from zero, as i32.
int32_t parse_weight(const std::string& s) {
    char* endp = nullptr;
    float x = strtof(s.c_str(), &endp);
    if (endp == s.c_str() || *endp != '\0') throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "cannot parse weight '" + s + "'"};
    float y = x * 1000.0f;
    return (int32_t)roundf(y);
}

static std::vector<std::string> split_whitespace(const std::string& l) {
    std::vector<std::string> out;
    std::istringstream is(l);
    std::string t;
    while (is >> t) out.push_back(t);
    return out;
}

// pattern.rs:18-35: the LAST 2-field line whose pvalue (f32) > threshold (f32) gives min_score.
bool parse_threshold_file(const std::string& filename, float pwm_threshold, int32_t* out) {
    std::ifstream f(filename);
    if (!f) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Could not open file " + filename};  // pattern.rs:115
    bool found = false;
    std::string line;
    while (std::getline(f, line)) {
        std::vector<std::string> x = split_whitespace(line);
        if (x.size() == 2) {
            int32_t weight = parse_weight(x[0]);
            float pvalue = strtof(x[1].c_str(), nullptr);
            if (pvalue > pwm_threshold) {
                *out = weight;
                found = true;
            }
        }
    }
    return found;
}

// pattern.rs:103-112
std::vector<Weight> reverse_complement(const std::vector<Weight>& w) {
    std::vector<Weight> x(w.rbegin(), w.rend());
    for (Weight& c : x) c = Weight::make(c.acgtn[3], c.acgtn[2], c.acgtn[1], c.acgtn[0]);
    return x;
}

// pattern.rs:89-101: first non-empty line is the name, rows of exactly 4 fields are columns A C G T.
static void parse_pwm_definition(const std::string& chunk, std::string* name, std::vector<Weight>* weights) {
    std::vector<std::string> lines;
    size_t p = 0;
    while (p <= chunk.size()) {
        size_t q = chunk.find('\n', p);
        if (q == std::string::npos) q = chunk.size();
        if (q > p) lines.push_back(chunk.substr(p, q - p));
        p = q + 1;
    }
    if (lines.empty()) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "empty PWM definition"};
    *name = lines[0];
    for (size_t i = 1; i < lines.size(); ++i) {
        std::vector<std::string> f = split_whitespace(lines[i]);
        if (f.size() == 4)
            weights->push_back(Weight::make(parse_weight(f[0]), parse_weight(f[1]), parse_weight(f[2]), parse_weight(f[3])));
    }
}

// pattern.rs:37-87: pattern_id counts wanted PWMs in FILE order, also those without a threshold.
std::vector<Pattern> parse_pwm_files(const std::string& pwm_file, const std::string& threshold_dir, float pwm_threshold,
                                     const std::vector<std::string>& wanted, bool add_reverse_patterns) {
    std::map<std::string, int32_t> thresholds;
    std::string dir = threshold_dir;
    while (!dir.empty() && dir.back() == '/') dir.pop_back();  // trim_end_matches("/")
    for (const std::string& p : wanted) {
        int32_t ms;
        if (parse_threshold_file(dir + "/" + p + ".thr", pwm_threshold, &ms)) thresholds[p] = ms;
    }
    std::ifstream f(pwm_file);
    if (!f) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Could not open file " + pwm_file};
    std::stringstream ss;
    ss << f.rdbuf();
    std::string content = ss.str();

    std::vector<Pattern> pwms;
    uint16_t pattern_id = 0;
    size_t p = 0;
    while (p <= content.size()) {
        size_t q = content.find('>', p);
        if (q == std::string::npos) q = content.size();
        std::string chunk = content.substr(p, q - p);
        p = q + 1;
        if (chunk.size() < 1) continue;
        std::string name;
        std::vector<Weight> weights;
        parse_pwm_definition(chunk, &name, &weights);
        if (std::find(wanted.begin(), wanted.end(), name) != wanted.end()) {
            auto it = thresholds.find(name);
            if (it != thresholds.end()) {
                Pattern fwd;
                fwd.weights = weights;
                fwd.name = name;
                fwd.pattern_id = pattern_id;
                fwd.min_score = it->second;
                fwd.direction = TFBS_DIR_P;
                pwms.push_back(fwd);
                if (add_reverse_patterns) {
                    Pattern rev = fwd;
                    rev.weights = reverse_complement(weights);
                    rev.direction = TFBS_DIR_N;
                    pwms.push_back(rev);
                }
            }
            pattern_id++;
        }
    }
    return pwms;
}

// pattern.rs:119-135: zip of columns and bases, N column = 0, i32 sum.
int32_t apply_pwm(const Pattern& p, const NucleotidePos* haplotype, size_t n) {
    if (!p.is_pwm) return 0;
    int32_t s = 0;
    size_t m = std::min(n, p.weights.size());
    for (size_t c = 0; c < m; ++c) s += p.weights[c].acgtn[haplotype[c].nuc];
    return s;
}

// pattern.rs:141-171
void matches(const Pattern& p, uint32_t pattern_index, const std::vector<NucleotidePos>& haplotype, uint32_t group,
             std::vector<Match>* out, uint64_t* cells) {
    if (!p.is_pwm) return;  // :166-168
    uint64_t p_length = pattern_length(p);
    if (haplotype.size() >= p.weights.size()) {  // :147
        size_t n = haplotype.size() - p.weights.size() + 1;
        if (cells) *cells += (uint64_t)n * p_length;
        for (size_t i = 0; i < n; ++i) {  // :149
            int32_t score = apply_pwm(p, &haplotype[i], haplotype.size() - i);
            if (score > p.min_score)  // :151 strict
                out->push_back(Match{Range{haplotype[i].pos, haplotype[i].pos + p_length - 1}, p.pattern_id, pattern_index, group});
        }
    }
}

// ---------------------------------------------------------------------------------------------
// main.rs
// ---------------------------------------------------------------------------------------------

// main.rs:94-154
RegionMatches find_all_matches(const Range& peak, const std::vector<std::vector<Diff>>& diffs_by_haplotype,
                               uint32_t variant_count, const std::vector<NucleotidePos>& ref_haplotype,
                               const std::vector<Pattern>& pwm_list, uint32_t sample_count) {
    RegionMatches rm;
    LoadedHaplotypes xs = load_haplotypes(peak, diffs_by_haplotype, variant_count, ref_haplotype);  // :99
    rm.number_of_variants = xs.variant_count;
    rm.sequence_collision = xs.sequence_collision;
    rm.overwritten_ids = xs.overwritten_ids;
    rm.truncated_ids = xs.truncated_ids;
    std::vector<uint8_t> has_reference(2 * (size_t)sample_count, 1);  // all_haplotype_ids :74-81
    rm.group_ids.emplace_back();                                      // group 0 = reference haplotype
    rm.group_len.push_back((uint32_t)ref_haplotype.size());
    for (size_t gi = 0; gi < xs.groups.size(); ++gi) {                // :101-126
        const HaplotypeGroup& g = xs.groups[gi];
        rm.number_of_haplotypes++;
        rm.truncated |= g.truncated;
        for (uint32_t h : g.haplotype_ids) has_reference[h] = 0;  // :103-105
        uint32_t gid = (uint32_t)rm.group_ids.size();
        rm.group_ids.push_back(g.haplotype_ids);
        rm.group_len.push_back((uint32_t)g.sequence.size());
        for (uint32_t pi = 0; pi < pwm_list.size(); ++pi)  // :113-124
            matches(pwm_list[pi], pi, g.sequence, gid, &rm.match_list, &rm.executed_cells);
    }
    for (uint32_t h = 0; h < has_reference.size(); ++h)
        if (has_reference[h]) rm.group_ids[0].push_back(h);
    if (!rm.group_ids[0].empty()) {  // :129-147
        rm.number_of_haplotypes++;
        for (uint32_t pi = 0; pi < pwm_list.size(); ++pi)
            matches(pwm_list[pi], pi, ref_haplotype, 0, &rm.match_list, &rm.executed_cells);
    }
    return rm;
}

// main.rs:62-72: p.overlaps(merged) is asymmetric: an original region strictly inside the merged
// one (touching neither end) is never selected (App. A.6 Q1).  Each occurrence is kept (Q3).
std::vector<InnerPeak> select_inner_peaks(const Range& peak, const std::vector<std::vector<Range>>& peak_map) {
    std::vector<InnerPeak> ip;
    for (uint32_t b = 0; b < peak_map.size(); ++b)
        for (const Range& p : peak_map[b])
            if (p.overlaps(peak)) ip.push_back(InnerPeak{b, p, 0});
    return ip;
}

// main.rs:500-534: entries are created on the first overlapping hit only; inner.overlaps(hit.range)
// is the asymmetric test (App. A.6 Q2); an identical range listed twice is visited twice (Q3).
std::map<CountKey, CountValue> count_matches_by_sample(const RegionMatches& rm, const std::vector<InnerPeak>& inner_peaks,
                                                       uint32_t sample_count) {
    std::map<CountKey, CountValue> pppp;
    for (const Match& m : rm.match_list) {
        for (const InnerPeak& ip : inner_peaks) {
            if (!ip.range.overlaps(m.range)) continue;
            CountKey key{ip.bed_index, ip.range, m.pattern_id};
            auto it = pppp.find(key);
            if (it == pppp.end()) {
                CountValue v;
                v.left.assign(sample_count, 0);
                v.right.assign(sample_count, 0);
                v.inner_index = ip.inner_index;
                it = pppp.emplace(key, std::move(v)).first;
            }
            for (uint32_t h : rm.group_ids[m.group]) {
                if ((h & 1) == 0) it->second.left[h >> 1] += 1;
                else it->second.right[h >> 1] += 1;
            }
        }
    }
    return pppp;
}

// main.rs:439-498
bool counts_as_genotypes(const std::vector<uint32_t>& v1, const std::vector<uint32_t>& v2, GenotypeRow* out) {
    std::vector<uint32_t> v(v1);
    for (size_t i = 0; i < v2.size(); ++i) v[i] += v2[i];
    if (v.empty()) return false;
    uint32_t lowest = *std::min_element(v.begin(), v.end());
    uint32_t highest = *std::max_element(v.begin(), v.end());
    if (lowest == highest) return false;  // :456-458
    std::string res;
    res.reserve(v.size() * 8);
    uint32_t intermediate_1_1000 = (lowest * 1000u * 3u + highest * 1000u) / 4u;  // :461 (u32, wrapping like release)
    uint32_t intermediate_3_1000 = (lowest * 1000u + highest * 1000u * 3u) / 4u;  // :462
    std::vector<uint32_t> all_values{lowest, highest};
    uint32_t zero_count = 0, one_count = 0, two_count = 0;
    float lowest_f32 = (float)lowest;
    float spread_f32 = (float)highest - lowest_f32;
    char buf[64];
    for (uint32_t x : v) {
        if (x == lowest) { res += "\t0|0:0.0"; zero_count++; }
        else if (x == highest) { res += "\t1|1:2.0"; two_count++; }
        else {
            if (std::find(all_values.begin(), all_values.end(), x) == all_values.end()) all_values.push_back(x);
            uint32_t x_1000 = x * 1000u;
            if (x_1000 < intermediate_1_1000) { res += "\t0|0"; zero_count++; }
            else if (x_1000 < intermediate_3_1000) { res += "\t0|1"; one_count++; }
            else { res += "\t1|1"; two_count++; }
            volatile float num = ((float)x - lowest_f32) * 2.0f;  // :478 f32 arithmetic, one rounding per op
            float pseudo_dosage = num / spread_f32;
            snprintf(buf, sizeof buf, ":%.4f", (double)pseudo_dosage);  // {:.4}: exact value, ties to even
            res += buf;
        }
    }
    uint32_t maf;  // :482-489
    if (zero_count >= one_count && zero_count >= two_count) maf = one_count + two_count;
    else if (two_count >= zero_count && two_count >= one_count) maf = zero_count + one_count;
    else maf = zero_count + two_count;
    std::sort(all_values.begin(), all_values.end());
    out->distinct_counts = all_values;
    out->maf = maf;
    out->freq0 = zero_count;
    out->freq1 = one_count;
    out->freq2 = two_count;
    out->genotypes = res;
    return true;
}

// ---------------------------------------------------------------------------------------------
// range.rs / bed.rs
// ---------------------------------------------------------------------------------------------

// range.rs:43-87: sort by start (stable), then fold with last.overlaps(range) -> merge.
std::vector<Range> range_stack(std::vector<Range> raw) {
    std::stable_sort(raw.begin(), raw.end(), [](const Range& a, const Range& b) { return a.start < b.start; });
    std::vector<Range> ranges;
    for (const Range& r : raw) {
        if (!ranges.empty() && ranges.back().overlaps(r)) ranges.back().merge(r);
        else ranges.push_back(r);
    }
    return ranges;
}

// bed.rs:9-19 via bio::io::bed::Reader: tab separated, '#' comment lines skipped, fields
// chrom,start,end; start/end used as an INCLUSIVE range.
std::vector<Range> load_bed(const std::string& filename, const std::string& chromosome) {
    std::ifstream f(filename);
    if (!f) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Bed file " + filename + " does not exist"};
    std::vector<Range> xs;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> fld;
        size_t p = 0;
        while (true) {
            size_t q = line.find('\t', p);
            if (q == std::string::npos) { fld.push_back(line.substr(p)); break; }
            fld.push_back(line.substr(p, q - p));
            p = q + 1;
        }
        if (fld.size() < 3) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "malformed BED line in " + filename + ": " + line};
        char* e1 = nullptr;
        char* e2 = nullptr;
        uint64_t s = strtoull(fld[1].c_str(), &e1, 10), e = strtoull(fld[2].c_str(), &e2, 10);
        if (*e1 || *e2 || fld[1].empty() || fld[2].empty())
            throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "malformed BED line in " + filename + ": " + line};
        if (fld[0] == chromosome) xs.push_back(Range{s, e});
    }
    return xs;
}

static std::string basename_of(const std::string& s) {  // bed.rs:49-52
    size_t p = s.find_last_of('/');
    return p == std::string::npos ? s : s.substr(p + 1);
}

// bed.rs:25-60.  peak_map is indexed by bed file (command-line order) instead of by basename;
// names carries the basenames.
void load_peak_files(const std::vector<std::string>& bed_files, const std::string& chromosome, uint64_t after_position,
                     std::vector<Range>* merged, std::vector<std::vector<Range>>* peak_map, std::vector<std::string>* bed_names) {
    peak_map->clear();
    bed_names->clear();
    std::vector<Range> all;
    for (const std::string& bf : bed_files) {
        std::vector<Range> peaks = load_bed(bf, chromosome);
        std::vector<Range> kept;
        for (const Range& p : peaks)
            if (p.start >= after_position) kept.push_back(p);  // :31
        all.insert(all.end(), kept.begin(), kept.end());
        peak_map->push_back(kept);
        bed_names->push_back(basename_of(bf));
    }
    *merged = range_stack(all);  // :37-43 (already sorted by start)
}

// ---------------------------------------------------------------------------------------------
// Block-level entry: the same structs the product's C ABI takes.
// ---------------------------------------------------------------------------------------------

std::vector<Pattern> patterns_from_c(const tfbs_pattern* p, uint32_t n) {
    std::vector<Pattern> out(n);
    for (uint32_t i = 0; i < n; ++i) {
        out[i].is_pwm = p[i].kind == TFBS_PATTERN_PWM;
        out[i].pattern_id = p[i].pattern_id;
        out[i].min_score = p[i].min_score;
        out[i].direction = p[i].direction;
        if (out[i].is_pwm)
            for (uint32_t c = 0; c < p[i].len; ++c)
                out[i].weights.push_back(Weight::make(p[i].weights[4 * c], p[i].weights[4 * c + 1], p[i].weights[4 * c + 2], p[i].weights[4 * c + 3]));
    }
    return out;
}

static uint64_t cells_for(size_t len, const std::vector<Pattern>& pwm_list) {
    uint64_t c = 0;
    for (const Pattern& p : pwm_list) {
        size_t L = pattern_length(p);
        if (p.is_pwm && len >= L) c += (uint64_t)(len - L + 1) * L;
    }
    return c;
}

void process_block_range(const std::vector<Pattern>& pwm_list, const tfbs_block& blk, uint32_t r0, uint32_t r1, int rows_mode,
                         bool want_matches, BlockResult* out) {
    const uint32_t S = blk.n_samples, H = 2 * S;
    for (uint32_t r = r0; r < r1; ++r) {
        if (blk.region_start[r] < 0 || blk.region_end[r] < blk.region_start[r])
            throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "region " + std::to_string(r) + " has an invalid extended window (main.rs:407 underflow)"};
        Range peak{(uint64_t)blk.region_start[r], (uint64_t)blk.region_end[r]};
        std::vector<NucleotidePos> ref_haplotype =
            to_nucleotides_pos(blk.ref_bases + blk.ref_off[r], (size_t)(blk.ref_off[r + 1] - blk.ref_off[r]), peak);  // main.rs:156-161

        // load_diffs, haplotype.rs:13-62 (GT decoding already folded into the carrier bits)
        std::vector<std::vector<Diff>> diffs_by_haplotype(H);
        uint32_t nvar = blk.var_off[r + 1] - blk.var_off[r];
        for (uint32_t vi = blk.var_off[r]; vi < blk.var_off[r + 1]; ++vi) {
            const tfbs_variant& v = blk.variants[vi];
            Diff d{(uint64_t)v.pos, to_nucleotides(blk.allele_bases + v.ref_off, v.ref_len),
                   to_nucleotides(blk.allele_bases + v.alt_off, v.alt_len)};
            const uint32_t* row = blk.carriers + (size_t)v.carrier_row * blk.carrier_pitch;
            for (uint32_t w = 0; w < (H + 31) / 32; ++w) {
                uint32_t bits = row[w];
                while (bits) {
                    uint32_t b = __builtin_ctz(bits);
                    bits &= bits - 1;
                    uint32_t h = 32 * w + b;
                    if (h < H) diffs_by_haplotype[h].push_back(d);
                }
            }
        }
        RegionMatches rm = find_all_matches(peak, diffs_by_haplotype, nvar, ref_haplotype, pwm_list, S);

        std::vector<InnerPeak> inner;
        for (uint32_t k = blk.inner_off[r]; k < blk.inner_off[r + 1]; ++k) {
            const tfbs_inner_region& ir = blk.inner[k];
            for (uint32_t m = 0; m < ir.multiplicity; ++m)  // Q3: visited once per occurrence
                inner.push_back(InnerPeak{ir.bed_index, Range{(uint64_t)ir.start, (uint64_t)ir.end}, k});
        }
        std::map<CountKey, CountValue> counts = count_matches_by_sample(rm, inner, S);

        std::vector<BlockRow> rows;
        for (auto& kv : counts) {
            uint32_t lo = UINT32_MAX, hi = 0;
            for (uint32_t s = 0; s < S; ++s) {
                uint32_t x = kv.second.left[s] + kv.second.right[s];
                lo = std::min(lo, x);
                hi = std::max(hi, x);
            }
            if (rows_mode == TFBS_ROWS_VARYING && lo == hi) continue;  // main.rs:456-458
            BlockRow br;
            br.region = r;
            br.inner = kv.second.inner_index;
            br.pattern_id = kv.first.pattern_id;
            br.vmin = lo;
            br.vmax = hi;
            br.left = std::move(kv.second.left);
            br.right = std::move(kv.second.right);
            rows.push_back(std::move(br));
        }
        std::sort(rows.begin(), rows.end(), [](const BlockRow& a, const BlockRow& b) {
            return a.pattern_id != b.pattern_id ? a.pattern_id < b.pattern_id : a.inner < b.inner;
        });
        for (BlockRow& br : rows) out->rows.push_back(std::move(br));

        out->executed_cells += rm.executed_cells;
        out->n_hits += rm.match_list.size();
        out->n_groups += rm.number_of_haplotypes;
        out->collision_regions += rm.sequence_collision ? 1 : 0;
        out->truncated_regions += rm.truncated ? 1 : 0;
        if (want_matches) {
            for (const Match& m : rm.match_list)
                out->matches.push_back(BlockMatch{r, m.pattern_index, m.group, (int64_t)m.range.start});
            size_t base = out->hap_group.size();
            out->hap_group.resize(base + H, 0);
            for (uint32_t g = 0; g < rm.group_ids.size(); ++g)
                for (uint32_t h : rm.group_ids[g]) out->hap_group[base + h] = g;
            out->hap_flags.resize(base + H, 0);
            for (uint32_t h : rm.truncated_ids) out->hap_flags[base + h] |= 1;
            for (uint32_t h : rm.overwritten_ids) out->hap_flags[base + h] |= 2;
        }
        // nominal = every haplotype of every sample scanned on its own sequence
        for (uint32_t g = 0; g < rm.group_ids.size(); ++g)
            out->nominal_cells += (uint64_t)rm.group_ids[g].size() * cells_for(rm.group_len[g], pwm_list);
    }
}

// main.rs:333-382: `threads` workers, each pulling chunks of 50 merged regions from a shared queue.
static std::atomic<uint32_t> g_chunk_size{50};
void set_chunk_size(uint32_t n) { g_chunk_size = n ? n : 50; }

void process_block(const std::vector<Pattern>& pwm_list, const tfbs_block& blk, int rows_mode, bool want_matches, int n_threads,
                   BlockResult* out) {
    const uint32_t CHUNK = g_chunk_size;  // 50 in the reference (main.rs:378); bench.py shrinks it for bounded samples
    uint32_t n_chunks = (blk.n_regions + CHUNK - 1) / CHUNK;
    std::vector<BlockResult> partial(n_chunks);
    std::atomic<uint32_t> next{0};
    std::mutex err_mu;
    OracleError first_err{0, ""};
    auto worker = [&]() {
        for (;;) {
            uint32_t c = next.fetch_add(1);
            if (c >= n_chunks) return;
            try {
                process_block_range(pwm_list, blk, c * CHUNK, std::min(blk.n_regions, (c + 1) * CHUNK), rows_mode, want_matches, &partial[c]);
            } catch (const OracleError& e) {
                std::lock_guard<std::mutex> g(err_mu);
                if (first_err.code == 0) first_err = e;
            }
        }
    };
    if (n_threads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (first_err.code != 0) throw first_err;
    for (BlockResult& p : partial) {
        for (BlockRow& r : p.rows) out->rows.push_back(std::move(r));
        out->matches.insert(out->matches.end(), p.matches.begin(), p.matches.end());
        out->hap_group.insert(out->hap_group.end(), p.hap_group.begin(), p.hap_group.end());
        out->hap_flags.insert(out->hap_flags.end(), p.hap_flags.begin(), p.hap_flags.end());
        out->executed_cells += p.executed_cells;
        out->nominal_cells += p.nominal_cells;
        out->n_groups += p.n_groups;
        out->n_hits += p.n_hits;
        out->collision_regions += p.collision_regions;
        out->truncated_regions += p.truncated_regions;
    }
}

}  // namespace ora
