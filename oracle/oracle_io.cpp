// oracle_io.cpp -- input decoding and the whole-program run() of the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// The reference reads its inputs through un-vendored third-party crates:
//   rust-htslib 0.26.1 (Cargo.lock:970-971)  BCF: haplotype.rs:16-24,78-79, main.rs:46-52
//   bio 0.28.2         (Cargo.lock:106-107)  FASTA: main.rs:156-159; BED: bed.rs:10-15
//   bgzip 0.0.3        (Cargo.lock:76-77)    output: main.rs:267
// Their published formats (BGZF, BCF2.2, .fai, BED) are restated here minimally and pinned
// by the reference's two integration fixtures (main.rs:548-568).
#include "oracle.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

namespace ora {

static std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Could not open file " + path};
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return d;
}

// BGZF = concatenated gzip members; zlib's gzip mode walks them one by one.
static std::vector<uint8_t> gunzip_all(const std::vector<uint8_t>& in, const std::string& what) {
    std::vector<uint8_t> out;
    size_t off = 0;
    while (off < in.size()) {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 16) != Z_OK) throw OracleError{TFBS_ERR_INTERNAL, "inflateInit2 failed"};
        zs.next_in = const_cast<Bytef*>(in.data() + off);
        zs.avail_in = (uInt)std::min<size_t>(in.size() - off, 1u << 30);
        int rc;
        do {
            size_t old = out.size();
            out.resize(old + (1u << 16));
            zs.next_out = out.data() + old;
            zs.avail_out = 1u << 16;
            rc = inflate(&zs, Z_NO_FLUSH);
            out.resize(old + ((1u << 16) - zs.avail_out));
            if (rc != Z_OK && rc != Z_STREAM_END) {
                inflateEnd(&zs);
                throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "corrupt gzip stream in " + what};
            }
        } while (rc != Z_STREAM_END);
        off += zs.total_in;
        inflateEnd(&zs);
    }
    return out;
}

std::string gunzip_file(const std::string& path) {
    std::vector<uint8_t> d = gunzip_all(read_file(path), path);
    return std::string(d.begin(), d.end());
}

namespace {
struct Cursor {
    const uint8_t* p;
    const uint8_t* e;
    void need(size_t n) const {
        if ((size_t)(e - p) < n) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "truncated BCF"};
    }
    uint8_t u8() { need(1); return *p++; }
    int32_t i32() { need(4); int32_t v; memcpy(&v, p, 4); p += 4; return v; }
    uint32_t u32() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
    int32_t typed_int_value(int type) {
        switch (type) {
            case 1: { need(1); int8_t v = (int8_t)*p; p += 1; return v; }
            case 2: { need(2); int16_t v; memcpy(&v, p, 2); p += 2; return v; }
            case 3: return i32();
            default: throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "BCF: integer expected"};
        }
    }
    // descriptor byte: low nibble type, high nibble length (15 = a typed int follows)
    void descriptor(int* type, uint32_t* len) {
        uint8_t b = u8();
        *type = b & 15;
        *len = b >> 4;
        if (*len == 15) {
            int t;
            uint32_t l;
            descriptor(&t, &l);
            *len = (uint32_t)typed_int_value(t);
        }
    }
    static size_t type_size(int type) {
        switch (type) {
            case 0: return 0;
            case 1: case 7: return 1;
            case 2: return 2;
            case 3: case 5: return 4;
            default: throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "BCF: unknown value type"};
        }
    }
    void skip_typed() {
        int t;
        uint32_t l;
        descriptor(&t, &l);
        size_t n = type_size(t) * l;
        need(n);
        p += n;
    }
    std::string typed_string() {
        int t;
        uint32_t l;
        descriptor(&t, &l);
        if (t != 7 && !(t == 0 && l == 0)) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "BCF: string expected"};
        need(l);
        std::string s((const char*)p, l);
        p += l;
        return s;
    }
};
}  // namespace

// BCF2.2: magic, l_text, header text, then records (l_shared, l_indiv, shared block, individual block).
// The dictionary of strings follows the header's FILTER/INFO/FORMAT lines (IDX= when present, else
// order of first appearance with PASS = 0); contigs likewise.
BcfFile read_bcf(const std::string& path) {
    std::vector<uint8_t> raw = gunzip_all(read_file(path), path);
    Cursor c{raw.data(), raw.data() + raw.size()};
    c.need(9);
    if (memcmp(c.p, "BCF\2", 4) != 0) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, path + " is not a BCF2 file"};
    c.p += 5;
    uint32_t l_text = c.u32();
    c.need(l_text);
    std::string text((const char*)c.p, l_text);
    c.p += l_text;

    BcfFile bf;
    std::vector<std::string> dict;  // string dictionary
    auto dict_set = [&](const std::string& id, int idx) {
        if (idx < 0) {
            if (std::find(dict.begin(), dict.end(), id) != dict.end()) return;
            dict.push_back(id);
        } else {
            if ((size_t)idx >= dict.size()) dict.resize(idx + 1);
            dict[idx] = id;
        }
    };
    dict_set("PASS", -1);
    std::istringstream is(text);
    std::string line;
    while (std::getline(is, line)) {
        while (!line.empty() && (line.back() == '\0' || line.back() == '\r')) line.pop_back();
        auto field = [&](const std::string& key) -> std::string {
            size_t p = line.find(key + "=");
            while (p != std::string::npos && p > 0 && line[p - 1] != '<' && line[p - 1] != ',') p = line.find(key + "=", p + 1);
            if (p == std::string::npos) return "";
            p += key.size() + 1;
            size_t q = line.find_first_of(",>", p);
            return line.substr(p, q == std::string::npos ? std::string::npos : q - p);
        };
        if (line.rfind("##contig=", 0) == 0) {
            std::string id = field("ID"), idx = field("IDX");
            if (!idx.empty()) {
                size_t k = (size_t)atoi(idx.c_str());
                if (k >= bf.contigs.size()) bf.contigs.resize(k + 1);
                bf.contigs[k] = id;
            } else bf.contigs.push_back(id);
        } else if (line.rfind("##FILTER=", 0) == 0 || line.rfind("##INFO=", 0) == 0 || line.rfind("##FORMAT=", 0) == 0) {
            std::string id = field("ID"), idx = field("IDX");
            dict_set(id, idx.empty() ? -1 : atoi(idx.c_str()));
        } else if (line.rfind("#CHROM", 0) == 0) {
            std::vector<std::string> f;
            size_t p = 0;
            for (;;) {
                size_t q = line.find('\t', p);
                f.push_back(line.substr(p, q == std::string::npos ? std::string::npos : q - p));
                if (q == std::string::npos) break;
                p = q + 1;
            }
            for (size_t i = 9; i < f.size(); ++i) bf.samples.push_back(f[i]);
        }
    }
    int gt_key = -1;
    for (size_t i = 0; i < dict.size(); ++i)
        if (dict[i] == "GT") gt_key = (int)i;

    while (c.p < c.e) {
        uint32_t l_shared = c.u32(), l_indiv = c.u32();
        c.need((size_t)l_shared + l_indiv);
        Cursor s{c.p, c.p + l_shared};
        Cursor d{c.p + l_shared, c.p + l_shared + l_indiv};
        c.p += (size_t)l_shared + l_indiv;
        BcfRecord r;
        r.rid = s.i32();
        r.pos = s.i32();
        r.rlen = s.i32();
        s.u32();  // QUAL
        uint32_t n_allele_info = s.u32(), n_fmt_sample = s.u32();
        uint32_t n_allele = n_allele_info >> 16, n_info = n_allele_info & 0xffff;
        uint32_t n_fmt = n_fmt_sample >> 24, n_sample = n_fmt_sample & 0xffffff;
        (void)n_info;
        s.typed_string();  // ID
        for (uint32_t a = 0; a < n_allele; ++a) r.alleles.push_back(s.typed_string());
        for (uint32_t f = 0; f < n_fmt; ++f) {
            int kt;
            uint32_t kl;
            d.descriptor(&kt, &kl);
            int32_t key = d.typed_int_value(kt);
            int vt;
            uint32_t vl;
            d.descriptor(&vt, &vl);
            size_t bytes = Cursor::type_size(vt) * vl * (size_t)n_sample;
            d.need(bytes);
            if (key == gt_key && vt >= 1 && vt <= 3) {
                r.gt_ploidy = (int)vl;
                r.gt.resize((size_t)vl * n_sample);
                Cursor g{d.p, d.p + bytes};
                for (size_t i = 0; i < r.gt.size(); ++i) r.gt[i] = g.typed_int_value(vt);
            }
            d.p += bytes;
        }
        bf.records.push_back(std::move(r));
    }
    return bf;
}

// bio::io::fasta::IndexedReader::fetch(chrom, start, stop) + read: bases [start, stop) of the
// contig through the .fai offsets (name, length, offset, line_bases, line_bytes).  A stop past the
// contig end is an error in bio 0.28's IndexedReader::read ("FASTA read interval was out of bounds"), which
// process_peak turns into a panic (main.rs:157-159); no reference test pins it (SURVEY 8c "unpinned").
std::vector<uint8_t> fasta_fetch(const std::string& fasta_path, const std::string& chrom, uint64_t start, uint64_t stop) {
    std::ifstream fai(fasta_path + ".fai");
    if (!fai) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Error while opening the reference genome index '" + fasta_path + ".fai'"};
    std::string name;
    uint64_t len = 0, offset = 0, line_bases = 0, line_bytes = 0;
    bool found = false;
    std::string line;
    while (std::getline(fai, line)) {
        std::istringstream ls(line);
        std::string n;
        uint64_t a, b, cc, dd;
        if (ls >> n >> a >> b >> cc >> dd && n == chrom) {
            len = a; offset = b; line_bases = cc; line_bytes = dd;
            found = true;
            break;
        }
    }
    if (!found) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Error while seeking in reference genome file: unknown sequence " + chrom};
    if (stop > len)
        throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Error while reading in reference genome file: FASTA read interval was out of bounds"};
    std::vector<uint8_t> out;
    if (start >= stop) return out;
    std::ifstream f(fasta_path, std::ios::binary);
    if (!f) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Error while opening the reference genome '" + fasta_path + "'"};
    out.reserve(stop - start);
    uint64_t pos = start;
    while (pos < stop) {
        uint64_t ln = pos / line_bases, col = pos % line_bases;
        uint64_t take = std::min(stop - pos, line_bases - col);
        f.seekg((std::streamoff)(offset + ln * line_bytes + col));
        size_t old = out.size();
        out.resize(old + take);
        f.read((char*)out.data() + old, (std::streamsize)take);
        if ((uint64_t)f.gcount() != take) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Error while reading in reference genome file"};
        pos += take;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// run(), main.rs:234-393, single worker; rows of a region are emitted in the oracle's canonical
// order (bed index, inner range, pattern_id) instead of HashMap::drain() order.
// ---------------------------------------------------------------------------------------------

// Owns the arrays a tfbs_block points into.
struct BlockStorage {
    std::vector<int64_t> region_start, region_end;
    std::vector<uint64_t> ref_off{0};
    std::vector<uint8_t> ref_bases;
    std::vector<uint32_t> inner_off{0};
    std::vector<tfbs_inner_region> inner;
    std::vector<uint32_t> var_off{0};
    std::vector<tfbs_variant> variants;
    std::vector<uint8_t> allele_bases;
    std::vector<uint32_t> carriers;
    uint32_t pitch = 0;
    tfbs_block view(uint32_t n_samples) const {
        tfbs_block b;
        memset(&b, 0, sizeof b);
        b.n_regions = (uint32_t)region_start.size();
        b.n_samples = n_samples;
        b.region_start = region_start.data();
        b.region_end = region_end.data();
        b.ref_off = ref_off.data();
        b.ref_bases = ref_bases.data();
        b.inner_off = inner_off.data();
        b.inner = inner.data();
        b.var_off = var_off.data();
        b.variants = variants.data();
        b.allele_bases = allele_bases.data();
        b.allele_bytes = allele_bases.size();
        b.carriers = carriers.data();
        b.n_carrier_rows = pitch ? (uint32_t)(carriers.size() / pitch) : 0;
        b.carrier_pitch = pitch;
        return b;
    }
};

std::string run(const RunOptions& opt) {
    std::vector<Pattern> pwm_list = parse_pwm_files(opt.pwm_file, opt.pwm_threshold_dir, opt.pwm_threshold, opt.wanted_pwms, !opt.forward_only);  // :237
    if (pwm_list.empty()) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "assertion failed: pwm_list.len() > 0"};          // :238
    std::map<uint16_t, std::string> pwm_name_dict;                                                                        // :239-250
    for (const Pattern& p : pwm_list) pwm_name_dict[p.pattern_id] = p.name;

    std::vector<Range> merged;
    std::vector<std::vector<Range>> peak_map;
    std::vector<std::string> bed_names;
    load_peak_files(opt.bed_files, opt.chromosome, opt.after_position, &merged, &peak_map, &bed_names);  // :252

    BcfFile bcf = read_bcf(opt.bcf);  // :255
    // :293-313 selected samples, always in BCF column order
    std::vector<std::string> samples;
    std::vector<size_t> sample_positions;
    if (!opt.has_samples) {
        samples = bcf.samples;
        for (size_t i = 0; i < bcf.samples.size(); ++i) sample_positions.push_back(i);
    } else {
        std::ifstream sf(opt.samples_file);
        if (!sf) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Could not open sample file " + opt.samples_file};
        std::set<std::string> wanted;
        std::string l;
        while (std::getline(sf, l)) {
            if (!l.empty() && l.back() == '\r') l.pop_back();
            if (l.size() > 1) wanted.insert(l);
        }
        for (size_t i = 0; i < bcf.samples.size(); ++i)
            if (wanted.count(bcf.samples[i])) {
                sample_positions.push_back(i);
                samples.push_back(bcf.samples[i]);
            }
    }
    const uint32_t S = (uint32_t)samples.size();

    std::string out = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";  // :320-324
    for (const std::string& s : samples) out += "\t" + s;
    out += "\n";

    int rid = -1;  // haplotype.rs:78 name2rid
    for (size_t i = 0; i < bcf.contigs.size(); ++i)
        if (bcf.contigs[i] == opt.chromosome) rid = (int)i;
    if (rid < 0) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "chromosome " + opt.chromosome + " is not in the BCF header"};

    uint32_t largest = 0;  // main.rs:404
    for (const Pattern& p : pwm_list) largest = std::max(largest, pattern_length(p));

    std::string chr = opt.chromosome;  // :402 replace("chr", "")
    for (size_t p; (p = chr.find("chr")) != std::string::npos;) chr.erase(p, 3);

    BlockStorage bs;
    bs.pitch = (2 * S + 31) / 32;
    if (bs.pitch == 0) bs.pitch = 1;
    std::vector<std::vector<InnerPeak>> inner_by_region;
    for (const Range& m : merged) {
        if (m.start + 1 < largest) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "attempt to subtract with overflow (main.rs:407)"};
        Range ext{m.start - largest + 1, m.end + largest - 1};  // :407
        bs.region_start.push_back((int64_t)ext.start);
        bs.region_end.push_back((int64_t)ext.end);
        std::vector<uint8_t> text = fasta_fetch(opt.reference, opt.chromosome, ext.start, ext.end + 1);  // :157
        bs.ref_bases.insert(bs.ref_bases.end(), text.begin(), text.end());
        bs.ref_off.push_back(bs.ref_bases.size());

        std::vector<InnerPeak> ip = select_inner_peaks(m, peak_map);  // :411
        // collapse identical (bed, range) occurrences into a multiplicity (A.6 Q3)
        for (const InnerPeak& p : ip) {
            bool dup = false;
            for (size_t k = bs.inner_off.back(); k < bs.inner.size(); ++k)
                if (bs.inner[k].bed_index == p.bed_index && (uint64_t)bs.inner[k].start == p.range.start && (uint64_t)bs.inner[k].end == p.range.end) {
                    bs.inner[k].multiplicity++;
                    dup = true;
                }
            if (!dup) bs.inner.push_back(tfbs_inner_region{(int64_t)p.range.start, (int64_t)p.range.end, p.bed_index, 1});
        }
        bs.inner_off.push_back((uint32_t)bs.inner.size());

        // reader.fetch(rid, start, end + 1) (haplotype.rs:79): records overlapping [start, end + 1)
        for (const BcfRecord& r : bcf.records) {
            if (r.rid != rid) continue;
            int64_t rbeg = r.pos, rend = r.pos + std::max<int32_t>(r.rlen, 1);
            if (!(rbeg < (int64_t)ext.end + 1 && rend > (int64_t)ext.start)) continue;
            if (r.alleles.size() < 2) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "index out of bounds: record without ALT (haplotype.rs:22)"};
            to_nucleotides((const uint8_t*)r.alleles[0].data(), r.alleles[0].size());  // :21-22 panic on unknown letters
            to_nucleotides((const uint8_t*)r.alleles[1].data(), r.alleles[1].size());
            if (r.alleles.size() != 2) continue;  // :27,53-55 skipped, only counted
            tfbs_variant v;
            memset(&v, 0, sizeof v);
            v.pos = r.pos;
            v.ref_off = (uint32_t)bs.allele_bases.size();
            v.ref_len = (uint32_t)r.alleles[0].size();
            bs.allele_bases.insert(bs.allele_bases.end(), r.alleles[0].begin(), r.alleles[0].end());
            v.alt_off = (uint32_t)bs.allele_bases.size();
            v.alt_len = (uint32_t)r.alleles[1].size();
            bs.allele_bases.insert(bs.allele_bases.end(), r.alleles[1].begin(), r.alleles[1].end());
            v.carrier_row = (uint32_t)(bs.carriers.size() / bs.pitch);
            bs.carriers.resize(bs.carriers.size() + bs.pitch, 0);
            uint32_t* row = bs.carriers.data() + (size_t)v.carrier_row * bs.pitch;
            for (uint32_t s = 0; s < S; ++s) {  // :30-51
                if (r.gt_ploidy != 2) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Inconsistent number of alleles"};  // :32
                int32_t g0 = r.gt[sample_positions[s] * 2], g1 = r.gt[sample_positions[s] * 2 + 1];
                if (g0 == 4) row[(2 * s) / 32] |= 1u << ((2 * s) % 32);          // Unphased(1)
                if (g1 == 5) row[(2 * s + 1) / 32] |= 1u << ((2 * s + 1) % 32);  // Phased(1)
            }
            bs.variants.push_back(v);
        }
        bs.var_off.push_back((uint32_t)bs.variants.size());
    }

    tfbs_block blk = bs.view(S);
    BlockResult res;
    process_block(pwm_list, blk, TFBS_ROWS_VARYING, false, (int)std::max<uint32_t>(1, opt.threads), &res);

    // rows arrive by region, then (pattern_id, inner); the canonical text order inside a region is
    // (bed index, inner range, pattern_id), POS is the running counter (:329,424-425).
    std::stable_sort(res.rows.begin(), res.rows.end(), [&](const BlockRow& a, const BlockRow& b) {
        if (a.region != b.region) return a.region < b.region;
        const tfbs_inner_region& ia = bs.inner[a.inner];
        const tfbs_inner_region& ib = bs.inner[b.inner];
        if (ia.bed_index != ib.bed_index) return ia.bed_index < ib.bed_index;
        if (ia.start != ib.start) return ia.start < ib.start;
        if (ia.end != ib.end) return ia.end < ib.end;
        return a.pattern_id < b.pattern_id;
    });
    uint32_t fake_position = 1;
    for (const BlockRow& row : res.rows) {
        GenotypeRow g;
        if (!counts_as_genotypes(row.left, row.right, &g)) continue;  // :420
        if (g.maf < opt.min_maf) continue;                             // :421
        const tfbs_inner_region& ir = bs.inner[row.inner];
        std::string info = "COUNTS=";
        for (size_t i = 0; i < g.distinct_counts.size(); ++i) info += (i ? "," : "") + std::to_string(g.distinct_counts[i]);
        info += ";freqs=" + std::to_string(g.freq0) + "/" + std::to_string(g.freq1) + "/" + std::to_string(g.freq2);
        out += chr + "\t" + std::to_string(fake_position++) + "\t" + bed_names[ir.bed_index] + "," + pwm_name_dict[row.pattern_id] + "," +
               std::to_string(ir.start) + "-" + std::to_string(ir.end) + "\t.\t.\t.\tPASS\t" + info + "\tGT:DS" + g.genotypes + "\n";
    }
    if (!opt.output.empty()) {
        if (opt.gzip_output) {
            gzFile gz = gzopen(opt.output.c_str(), "wb");
            if (!gz) throw OracleError{TFBS_ERR_INVALID_ARGUMENT, "Could not create output file"};
            gzwrite(gz, out.data(), (unsigned)out.size());
            gzclose(gz);
        } else {
            std::ofstream of(opt.output, std::ios::binary);
            of << out;
        }
    }
    return out;
}

}  // namespace ora
