// oracle.hpp -- CPU restatement of the find-tfbs hot path.  TEST INFRASTRUCTURE ONLY.
//
// This is the checker the CUDA path is compared against.  Nothing under find_tfbs_b200/
// (the product) includes, links or calls it; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py do.
//
// The Rust reference cannot be compiled in this image (no cargo/rustc/clang, no htslib), so
// this file restates its algorithm function by function, each citing the reference
// file:line it follows (paths relative to the reference repository).  It is pinned against
// every unit-test vector of the reference that touches this path and against both golden
// VCFs (tests/test_oracle_golden.py); what stays unpinned is listed in DESIGN.md.
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../include/tfbs.h"

namespace ora {

// src/types.rs:5-8
enum Nucleotide : uint8_t { A = 0, C = 1, G = 2, T = 3, N = 4 };

// src/types.rs:23-27
struct NucleotidePos {
    Nucleotide nuc;
    uint64_t pos;
    bool operator==(const NucleotidePos& o) const { return nuc == o.nuc && pos == o.pos; }
    bool operator<(const NucleotidePos& o) const { return nuc != o.nuc ? nuc < o.nuc : pos < o.pos; }
};

// src/range.rs:4-34
struct Range {
    uint64_t start, end;
    bool overlaps(const Range& other) const {  // asymmetric on purpose, range.rs:18-21
        return (other.start >= start && other.start <= end) || (other.end >= start && other.end <= end);
    }
    bool contains(uint64_t p) const { return p >= start && p <= end; }
    void merge(const Range& o) {
        if (o.start < start) start = o.start;
        if (o.end > end) end = o.end;
    }
    bool operator==(const Range& o) const { return start == o.start && end == o.end; }
    bool operator<(const Range& o) const { return start != o.start ? start < o.start : end < o.end; }
};

// src/types.rs:39-44; the derived Ord is (pos, reference, alternative), vectors lexicographic.
struct Diff {
    uint64_t pos;
    std::vector<Nucleotide> reference;
    std::vector<Nucleotide> alternative;
    bool operator==(const Diff& o) const { return pos == o.pos && reference == o.reference && alternative == o.alternative; }
    bool operator<(const Diff& o) const {
        if (pos != o.pos) return pos < o.pos;
        if (reference != o.reference) return reference < o.reference;
        return alternative < o.alternative;
    }
};

// src/types.rs:66-70.  side 0 = Left, 1 = Right.
struct HaplotypeId {
    uint32_t sample_id;
    uint8_t side;
    uint32_t index() const { return 2 * sample_id + side; }
};

// src/types.rs:103-113
struct Weight {
    int32_t acgtn[5];
    static Weight make(int32_t a, int32_t c, int32_t g, int32_t t) { return Weight{{a, c, g, t, 0}}; }
    bool operator==(const Weight& o) const {
        for (int i = 0; i < 5; ++i)
            if (acgtn[i] != o.acgtn[i]) return false;
        return true;
    }
};

// src/types.rs:86-90
struct Pattern {
    bool is_pwm = true;
    std::vector<Weight> weights;
    std::string name;
    uint16_t pattern_id = 0;
    int32_t min_score = 0;
    uint8_t direction = TFBS_DIR_P;
};

inline uint32_t pattern_length(const Pattern& p) { return p.is_pwm ? (uint32_t)p.weights.size() : 0; }  // types.rs:92-101

// src/types.rs:32-37 (haplotype_ids is an index into the group table of the region)
struct Match {
    Range range;
    uint16_t pattern_id;
    uint32_t pattern_index;  // position in pwm_list (not in the reference; for hit-list parity)
    uint32_t group;          // which distinct haplotype produced it
};

struct OracleError {
    int code;
    std::string message;
};

// ---- util.rs ----------------------------------------------------------------------------
Nucleotide to_nucleotide(uint8_t l);                                                     // util.rs:4-16
std::vector<Nucleotide> to_nucleotides(const uint8_t* letters, size_t n);                // util.rs:18-20
std::vector<NucleotidePos> to_nucleotides_pos(const uint8_t* letters, size_t n, const Range& r);  // util.rs:22-31

// ---- haplotype.rs -----------------------------------------------------------------------
std::vector<NucleotidePos> patch_haplotype(const Range& range, const std::vector<Diff>& diffs,
                                           const std::vector<NucleotidePos>& ref_haplotype,
                                           bool* truncated = nullptr);                    // haplotype.rs:94-156

// One distinct haplotype of a region after load_haplotypes (haplotype.rs:77-88).
struct HaplotypeGroup {
    std::vector<NucleotidePos> sequence;
    std::vector<uint32_t> haplotype_ids;  // 2*sample+side, ascending
    std::vector<Diff> diffs;
    bool truncated = false;
};

struct LoadedHaplotypes {
    uint32_t variant_count = 0;
    std::vector<HaplotypeGroup> groups;      // survivors of the sequence-keyed map, by ascending first id
    std::vector<uint32_t> overwritten_ids;   // haplotypes that fell back to the reference (App. A.6 Q4)
    std::vector<uint32_t> truncated_ids;     // haplotypes whose diff list ran into the truncation exit (haplotype.rs:144-149)
    bool sequence_collision = false;         // two diff lists patched to the same non-reference sequence
};

// diffs_by_haplotype[h] = Vec<Diff> in record order for haplotype h (load_diffs, haplotype.rs:13-62).
LoadedHaplotypes load_haplotypes(const Range& peak, const std::vector<std::vector<Diff>>& diffs_by_haplotype,
                                 uint32_t variant_count, const std::vector<NucleotidePos>& ref_haplotype);

// ---- pattern.rs -------------------------------------------------------------------------
int32_t parse_weight(const std::string& s);                                              // pattern.rs:13-16
bool parse_threshold_file(const std::string& filename, float pwm_threshold, int32_t* out);  // pattern.rs:18-35
std::vector<Pattern> parse_pwm_files(const std::string& pwm_file, const std::string& threshold_dir,
                                     float pwm_threshold, const std::vector<std::string>& wanted,
                                     bool add_reverse_patterns);                         // pattern.rs:37-87
std::vector<Weight> reverse_complement(const std::vector<Weight>& w);                   // pattern.rs:103-112
int32_t apply_pwm(const Pattern& p, const NucleotidePos* haplotype, size_t n);           // pattern.rs:125-135
void matches(const Pattern& p, uint32_t pattern_index, const std::vector<NucleotidePos>& haplotype, uint32_t group,
             std::vector<Match>* out, uint64_t* cells = nullptr);                        // pattern.rs:141-171

// ---- main.rs ----------------------------------------------------------------------------
struct RegionMatches {
    std::vector<Match> match_list;
    // group table: group 0 = reference haplotype (ids = everybody not in another group), then
    // LoadedHaplotypes.groups in order.  group_ids[g] = haplotype indices.
    std::vector<std::vector<uint32_t>> group_ids;
    std::vector<uint32_t> group_len;  // sequence length of each group
    uint32_t number_of_haplotypes = 0;
    uint32_t number_of_variants = 0;
    uint64_t executed_cells = 0;
    bool sequence_collision = false;
    bool truncated = false;
    std::vector<uint32_t> overwritten_ids, truncated_ids;  // audit lists, see LoadedHaplotypes
};
RegionMatches find_all_matches(const Range& peak, const std::vector<std::vector<Diff>>& diffs_by_haplotype,
                               uint32_t variant_count, const std::vector<NucleotidePos>& ref_haplotype,
                               const std::vector<Pattern>& pwm_list, uint32_t sample_count);   // main.rs:94-154

// inner_peaks: (bed index, range) in visiting order; duplicates appear twice (A.6 Q3).
struct InnerPeak {
    uint32_t bed_index;
    Range range;
    uint32_t inner_index;  // caller-defined tag carried into the rows
};
struct CountKey {
    uint32_t bed_index;
    Range range;
    uint16_t pattern_id;
    bool operator<(const CountKey& o) const {
        if (bed_index != o.bed_index) return bed_index < o.bed_index;
        if (!(range == o.range)) return range < o.range;
        return pattern_id < o.pattern_id;
    }
};
struct CountValue {
    std::vector<uint32_t> left, right;
    uint32_t inner_index;
};
std::map<CountKey, CountValue> count_matches_by_sample(const RegionMatches& rm, const std::vector<InnerPeak>& inner_peaks,
                                                       uint32_t sample_count);          // main.rs:500-534

struct GenotypeRow {
    std::vector<uint32_t> distinct_counts;
    uint32_t maf, freq0, freq1, freq2;
    std::string genotypes;
};
bool counts_as_genotypes(const std::vector<uint32_t>& v1, const std::vector<uint32_t>& v2, GenotypeRow* out);  // main.rs:439-498

std::vector<InnerPeak> select_inner_peaks(const Range& peak, const std::vector<std::vector<Range>>& peak_map);  // main.rs:62-72

// ---- bed.rs / range.rs ------------------------------------------------------------------
std::vector<Range> range_stack(std::vector<Range> raw);                                  // range.rs:43-87
std::vector<Range> load_bed(const std::string& filename, const std::string& chromosome); // bed.rs:9-19
void load_peak_files(const std::vector<std::string>& bed_files, const std::string& chromosome, uint64_t after_position,
                     std::vector<Range>* merged, std::vector<std::vector<Range>>* peak_map,
                     std::vector<std::string>* bed_names);                               // bed.rs:25-60

// ---- third-party edge (rust-htslib / bio), restated minimally ---------------------------
struct BcfRecord {
    int32_t rid;
    int64_t pos;
    int32_t rlen;
    std::vector<std::string> alleles;
    int gt_ploidy = 0;                 // values per sample in FORMAT/GT, 0 if absent
    std::vector<int32_t> gt;           // n_sample * gt_ploidy raw BCF codes
};
struct BcfFile {
    std::vector<std::string> contigs;
    std::vector<std::string> samples;
    std::vector<BcfRecord> records;    // whole file, file order
};
BcfFile read_bcf(const std::string& path);
std::vector<uint8_t> fasta_fetch(const std::string& fasta_path, const std::string& chrom, uint64_t start, uint64_t stop);  // [start, stop)
std::string gunzip_file(const std::string& path);  // all gzip members concatenated (BGZF)

// ---- block-level entry (same structs as the product's C ABI) ----------------------------
struct BlockRow {
    uint32_t region, inner;
    uint16_t pattern_id;
    uint32_t vmin, vmax;
    std::vector<uint32_t> left, right;
};
struct BlockMatch {
    uint32_t region, pattern_index, group;
    int64_t start;
};
struct BlockResult {
    std::vector<BlockRow> rows;
    std::vector<BlockMatch> matches;           // if requested
    std::vector<uint32_t> hap_group;           // [n_regions * 2S], if requested
    std::vector<uint8_t> hap_flags;            // [n_regions * 2S], if requested: bit0 truncated, bit1 overwritten (-> reference)
    uint64_t executed_cells = 0, nominal_cells = 0, n_groups = 0, n_hits = 0;
    uint32_t collision_regions = 0, truncated_regions = 0;
};
std::vector<Pattern> patterns_from_c(const tfbs_pattern* p, uint32_t n);
// Process regions [r0, r1) of a block.  Throws OracleError on the reference's panic conditions.
void process_block_range(const std::vector<Pattern>& pwm_list, const tfbs_block& blk, uint32_t r0, uint32_t r1,
                         int rows_mode, bool want_matches, BlockResult* out);
// Multi-threaded like the reference: n_threads workers pull chunks of 50 regions (main.rs:333-382).
void process_block(const std::vector<Pattern>& pwm_list, const tfbs_block& blk, int rows_mode, bool want_matches,
                   int n_threads, BlockResult* out);
// Work-queue granularity (regions per chunk); the reference uses 50 because BCF is block-compressed (main.rs:375-378).
void set_chunk_size(uint32_t n);

// ---- whole program (run(), main.rs:234-393) ---------------------------------------------
struct RunOptions {
    std::string chromosome, bcf, reference, pwm_file, pwm_threshold_dir, output, samples_file;
    std::vector<std::string> bed_files, wanted_pwms;
    float pwm_threshold = 0.f;
    bool forward_only = false, has_samples = false, verbose = false, gzip_output = true;
    uint32_t min_maf = 0, threads = 1;
    uint64_t after_position = 0;
};
// Returns the decompressed VCF text (also written to opt.output unless empty).
std::string run(const RunOptions& opt);

}  // namespace ora
