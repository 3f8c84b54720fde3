// bcf.hpp -- BCF2 decoding: records, GT -> carrier bits (haplotype.rs:13-62), CSI index
// Host side of find-tfbs-b200 (see driver.cpp for the map); header-only, one translation unit.
#pragma once
#include "options.hpp"
#include "bgzf.hpp"
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------------------------
// BCF2 (what rust-htslib's IndexedReader gives load_diffs: pos, alleles, GT of the selected samples)
// ---------------------------------------------------------------------------------------------------------------
struct Cursor {
    const uint8_t* p;
    const uint8_t* e;
    void need(size_t n) const { if ((size_t)(e - p) < n) die("truncated BCF"); }
    uint8_t u8() { need(1); return *p++; }
    int32_t i32() { need(4); int32_t v; memcpy(&v, p, 4); p += 4; return v; }
    uint32_t u32() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
    int32_t tint(int t) {
        if (t == 1) { need(1); return (int8_t)*p++; }
        if (t == 2) { need(2); int16_t v; memcpy(&v, p, 2); p += 2; return v; }
        if (t == 3) return i32();
        die("BCF: integer expected");
    }
    void desc(int* t, uint32_t* n) {
        uint8_t b = u8();
        *t = b & 15;
        *n = b >> 4;
        if (*n == 15) { int t2; uint32_t n2; desc(&t2, &n2); *n = (uint32_t)tint(t2); }
    }
    static size_t tsize(int t) {
        switch (t) { case 0: return 0; case 1: case 7: return 1; case 2: return 2; case 3: case 5: return 4; }
        die("BCF: unknown value type");
    }
    std::string tstr() {
        int t; uint32_t n;
        desc(&t, &n);
        if (t != 7 && !(t == 0 && n == 0)) die("BCF: string expected");
        need(n);
        std::string s((const char*)p, n);
        p += n;
        return s;
    }
};

struct Record {
    int64_t pos;
    int32_t rlen;
    uint32_t n_allele;
    std::string ref, alt;   // alleles[0], alleles[1]
    uint32_t carrier_row;   // row in Cohort::carriers (biallelic records only), else UINT32_MAX
    // What the reference would panic with when this record is fetched for a region (haplotype.rs:21-32: alleles[1] of a record
    // without ALT, a letter outside ACGTN in REF / ALT, a missing GT, a ploidy other than 2).  The reference only ever looks at
    // records inside an extended peak, so the message is kept here and raised by build_block for such records only.
    std::string problem;
};

struct Cohort {
    std::vector<std::string> bcf_samples, samples;  // all columns / selected, in BCF order
    std::vector<size_t> sample_positions;
    std::vector<Record> records;                    // of the wanted chromosome, file order (sorted by pos)
    std::vector<uint32_t> carriers;                 // [rows][pitch]
    uint32_t pitch = 1;
    int32_t max_rlen = 1;
};

void parse_bcf_header(Cursor& c, std::vector<std::string>* contigs, std::vector<std::string>* samples, int* gt_key) {
    c.need(9);
    if (memcmp(c.p, "BCF\2", 4) != 0) die("Error while opening the bcf file: not a BCF2 file");
    c.p += 5;
    uint32_t l_text = c.u32();
    c.need(l_text);
    std::string text((const char*)c.p, l_text);
    c.p += l_text;
    std::vector<std::string> dict{"PASS"};
    auto dict_set = [&](const std::string& id, int idx) {
        if (idx < 0) { if (std::find(dict.begin(), dict.end(), id) == dict.end()) dict.push_back(id); }
        else { if ((size_t)idx >= dict.size()) dict.resize(idx + 1); dict[idx] = id; }
    };
    for (std::string line : split(text, '\n')) {
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
            if (!idx.empty()) { size_t k = (size_t)atoi(idx.c_str()); if (k >= contigs->size()) contigs->resize(k + 1); (*contigs)[k] = id; }
            else contigs->push_back(id);
        } else if (line.rfind("##FILTER=", 0) == 0 || line.rfind("##INFO=", 0) == 0 || line.rfind("##FORMAT=", 0) == 0) {
            std::string id = field("ID"), idx = field("IDX");
            dict_set(id, idx.empty() ? -1 : atoi(idx.c_str()));
        } else if (line.rfind("#CHROM", 0) == 0) {
            auto f = split(line, '\t');
            for (size_t i = 9; i < f.size(); ++i) samples->push_back(f[i]);
        }
    }
    *gt_key = -1;
    for (size_t i = 0; i < dict.size(); ++i)
        if (dict[i] == "GT") *gt_key = (int)i;
}

std::string check_letters(const std::string& s) {  // util.rs:4-16; "" = fine
    for (unsigned char l : s)
        if (!(l == 65 || l == 97 || l == 67 || l == 99 || l == 71 || l == 103 || l == 84 || l == 116 || l == 78 || l == 110))
            return "Unknown nucleotide " + std::to_string((int)l);
    return "";
}

// Where the records of contig `rid` live in the BGZF file, from the CSI index next to the BCF (<bcf>.csi).  The reference reads
// through htslib's IndexedReader, which seeks with the same index (haplotype.rs:78-79); here the whole contig is taken at once.
// Virtual offsets are (compressed offset of the member << 16) | offset inside the inflated member.
struct CsiSpan {
    bool found = false;   // an index was read and holds this contig
    bool empty = false;   // ... and the contig has no record
    uint64_t vbeg = 0, vend = 0;
};
CsiSpan csi_contig_span(const std::string& bcf_path, int rid) {
    CsiSpan sp;
    std::ifstream probe(bcf_path + ".csi", std::ios::binary);
    if (!probe) return sp;
    probe.close();
    MappedFile f(bcf_path + ".csi", "Error while opening the index");
    const RawBytes d = gunzip_bgzf(f.data(), f.size(), bcf_path + ".csi", 1);
    Cursor c{d.data(), d.data() + d.size()};
    auto u64 = [&] { c.need(8); uint64_t v; memcpy(&v, c.p, 8); c.p += 8; return v; };
    if (d.size() < 16 || memcmp(c.p, "CSI\1", 4) != 0) return sp;
    c.p += 4;
    c.i32();  // min_shift
    const int32_t depth = c.i32();
    const int32_t l_aux = c.i32();
    if (l_aux < 0 || depth < 0 || depth > 10) return sp;
    c.need((size_t)l_aux);
    c.p += l_aux;
    const int32_t n_ref = c.i32();
    if (rid < 0 || rid >= n_ref) return sp;
    const uint32_t pseudo_bin = (uint32_t)((((uint64_t)1 << (3 * depth + 3)) - 1) / 7 + 1);  // holds statistics, not records
    for (int32_t r = 0; r <= rid; ++r) {
        const int32_t n_bin = c.i32();
        for (int32_t b = 0; b < n_bin; ++b) {
            const uint32_t bin = c.u32();
            u64();  // loffset
            const int32_t n_chunk = c.i32();
            for (int32_t k = 0; k < n_chunk; ++k) {
                const uint64_t beg = u64(), end = u64();
                if (r != rid || bin == pseudo_bin) continue;
                if (!sp.found || beg < sp.vbeg) sp.vbeg = beg;
                if (!sp.found || end > sp.vend) sp.vend = end;
                sp.found = true;
            }
        }
    }
    if (!sp.found) { sp.found = true; sp.empty = true; }
    return sp;
}

Cohort load_bcf(const Options& o) {
    const bool timing = getenv("TFBS_DRIVER_TIMING") != nullptr;  // stderr: where the time of the BCF phase goes
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "load_bcf: %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_mark).count());
        t_mark = now;
    };
    MappedFile file(o.bcf, "Error while opening the bcf file");
    Cohort co;
    std::vector<std::string> contigs;
    int gt_key;
    size_t header_bytes = 0;
    {   // the header sits in the first members: inflate only as many as it needs
        std::vector<uint8_t> head = gunzip_members(file.data(), file.size(), o.bcf, 9);
        if (head.size() >= 9 && memcmp(head.data(), "BCF\2", 4) == 0) {
            uint32_t l_text;
            memcpy(&l_text, head.data() + 5, 4);
            header_bytes = 9 + (size_t)l_text;
            if (head.size() < header_bytes) head = gunzip_members(file.data(), file.size(), o.bcf, header_bytes);
        }
        Cursor hc{head.data(), head.data() + head.size()};
        parse_bcf_header(hc, &contigs, &co.bcf_samples, &gt_key);
    }
    // main.rs:293-313: the selection is always in BCF column order
    if (!o.has_samples) {
        co.samples = co.bcf_samples;
        for (size_t i = 0; i < co.bcf_samples.size(); ++i) co.sample_positions.push_back(i);
    } else {
        std::ifstream sf(o.samples_file);
        if (!sf) die("Could not open sample file " + o.samples_file);
        std::set<std::string> wanted;
        std::string l;
        while (std::getline(sf, l)) {
            if (!l.empty() && l.back() == '\r') l.pop_back();
            if (l.size() > 1) wanted.insert(l);
        }
        for (size_t i = 0; i < co.bcf_samples.size(); ++i)
            if (wanted.count(co.bcf_samples[i])) { co.sample_positions.push_back(i); co.samples.push_back(co.bcf_samples[i]); }
    }
    printf("Reading %zu samples out of %zu\n", co.samples.size(), co.bcf_samples.size());
    int rid = -1;
    for (size_t i = 0; i < contigs.size(); ++i)
        if (contigs[i] == o.chromosome) rid = (int)i;
    if (rid < 0) die("called `Result::unwrap()` on an `Err` value: UnknownSequence (" + o.chromosome + ")");  // haplotype.rs:78
    const uint32_t S = (uint32_t)co.samples.size();
    co.pitch = std::max<uint32_t>(1, (2 * S + 31) / 32);
    // With a CSI index only the BGZF members that hold the wanted contig are read and inflated; without one the whole file is.
    RawBytes raw;
    size_t first_record = header_bytes;
    const CsiSpan span = o.use_index ? csi_contig_span(o.bcf, rid) : CsiSpan();
    const bool indexed = span.found && (span.empty || ((span.vbeg >> 16) < file.size() && bgzf_member_size(file.data(), file.size(), span.vbeg >> 16)));
    if (indexed && !span.empty) {
        const size_t cbeg = (size_t)(span.vbeg >> 16);
        size_t cend = std::min<size_t>(file.size(), (size_t)(span.vend >> 16));
        if ((span.vend & 0xffff) && cend < file.size()) cend += bgzf_member_size(file.data(), file.size(), cend);  // the last member is used in part
        if (cend <= cbeg) cend = file.size();
        raw = gunzip_bgzf(file.data() + cbeg, cend - cbeg, o.bcf, std::max(1u, o.threads));
        first_record = (size_t)(span.vbeg & 0xffff);
    } else if (!indexed) {
        raw = gunzip_bgzf(file.data(), file.size(), o.bcf, std::max(1u, o.threads));
    }
    lap("header + inflate");
    if (first_record > raw.size()) die("truncated BCF");
    Cursor c{raw.data() + first_record, raw.data() + raw.size()};
    // pass 1 (serial, a few bytes per record): positions, alleles, carrier rows; the genotype blocks are only located
    struct Pending { uint32_t row; const uint8_t* indiv; uint32_t l_indiv, n_fmt, n_sample; size_t rec; };
    std::vector<Pending> pending;
    while (c.p < c.e) {
        if (indexed) {
            // the last member of the contig's span may end inside the first record of the next contig: an incomplete record there
            // is the end of the contig, not a truncated file
            uint32_t ls = 0, li = 0;
            if (c.e - c.p >= 8) { memcpy(&ls, c.p, 4); memcpy(&li, c.p + 4, 4); }
            if (c.e - c.p < 8 || (size_t)(c.e - c.p) - 8 < (size_t)ls + li) {
                if (c.e - c.p >= 12) { int32_t chrom_peek; memcpy(&chrom_peek, c.p + 8, 4); if (chrom_peek == rid) die("truncated BCF"); }
                break;
            }
        }
        uint32_t l_shared = c.u32(), l_indiv = c.u32();
        c.need((size_t)l_shared + l_indiv);
        Cursor s{c.p, c.p + l_shared};
        const uint8_t* indiv = c.p + l_shared;
        c.p += (size_t)l_shared + l_indiv;
        int32_t chrom = s.i32();
        Record r;
        r.pos = s.i32();
        r.rlen = s.i32();
        s.u32();
        uint32_t nai = s.u32(), nfs = s.u32();
        r.n_allele = nai >> 16;
        uint32_t n_fmt = nfs >> 24, n_sample = nfs & 0xffffff;
        if (chrom != rid) {
            if (indexed) break;  // the index pointed at the contig's first record: its records end here (the file is sorted)
            continue;
        }
        s.tstr();
        if (r.n_allele >= 1) r.ref = s.tstr();
        if (r.n_allele >= 2) r.alt = s.tstr();
        if (r.n_allele < 2) r.problem = "index out of bounds: the len is " + std::to_string(r.n_allele) + " but the index is 1";  // haplotype.rs:22
        if (r.problem.empty()) r.problem = check_letters(r.ref);  // haplotype.rs:21-22 convert alleles[0] and [1] of every fetched record
        if (r.problem.empty()) r.problem = check_letters(r.alt);
        r.carrier_row = UINT32_MAX;
        if (r.n_allele == 2 && r.problem.empty()) {
            r.carrier_row = (uint32_t)pending.size();
            pending.push_back(Pending{r.carrier_row, indiv, l_indiv, n_fmt, n_sample, co.records.size()});
        }
        co.max_rlen = std::max(co.max_rlen, std::max(1, r.rlen));
        co.records.push_back(std::move(r));
    }
    lap("pass 1 (records, serial)");
    bool all_samples_in_order = true;
    for (size_t k = 0; k < co.sample_positions.size(); ++k) all_samples_in_order = all_samples_in_order && co.sample_positions[k] == k;
    // pass 2 (--threads host threads): GT of the selected samples -> carrier bits (haplotype.rs:30-51), the O(records x samples) part
    co.carriers.assign(pending.size() * (size_t)co.pitch, 0);
    std::atomic<size_t> next_rec{0};
    std::mutex err_mu;
    std::string err;
    auto decode = [&] {
        try {
            for (;;) {
                const size_t base = next_rec.fetch_add(256);
                if (base >= pending.size()) return;
                for (size_t k2 = base; k2 < std::min(pending.size(), base + 256); ++k2) {
                    const Pending& pd = pending[k2];
                    Cursor d{pd.indiv, pd.indiv + pd.l_indiv};
                    uint32_t* row = co.carriers.data() + (size_t)pd.row * co.pitch;
                    bool have_gt = false;
                    std::string& problem = co.records[pd.rec].problem;  // one thread per record: no race
                    for (uint32_t f = 0; f < pd.n_fmt; ++f) {
                        int kt, vt; uint32_t kl, vl;
                        d.desc(&kt, &kl);
                        int32_t key = d.tint(kt);
                        d.desc(&vt, &vl);
                        size_t bytes = Cursor::tsize(vt) * vl * (size_t)pd.n_sample;
                        d.need(bytes);
                        if (key == gt_key && vt >= 1 && vt <= 3) {
                            have_gt = true;
                            if (vl != 2 && S) { problem = "Inconsistent number of alleles"; d.p += bytes; continue; }  // haplotype.rs:32
                            const size_t es = Cursor::tsize(vt);
                            if (vt == 1) {
                                // 8-bit GT vectors (every cohort with fewer than 63 alleles per record): the same rule without the
                                // generic cursor.  Bit 2k of the row = (first value of sample k == 4 = Unphased(1)), bit 2k + 1 =
                                // (second value == 5 = Phased(1)), haplotype.rs:34-41.
                                const int8_t* gp = (const int8_t*)d.p;
                                bool vend_seen = false;
                                if (all_samples_in_order) {
                                    // every sample, in file order: the row is the byte-wise comparison of the GT vector with 04 05 04 05 ...
                                    if ((size_t)S * 2 > bytes) die("truncated BCF");
                                    const size_t nb = (size_t)S * 2;
                                    size_t j = 0;
#if defined(__SSE2__)
                                    const __m128i pat = _mm_set1_epi16(0x0504), vend = _mm_set1_epi8((char)-127);
                                    __m128i any_vend = _mm_setzero_si128();
                                    for (; j + 32 <= nb; j += 32) {
                                        const __m128i a = _mm_loadu_si128((const __m128i*)(gp + j)), b = _mm_loadu_si128((const __m128i*)(gp + j + 16));
                                        any_vend = _mm_or_si128(any_vend, _mm_or_si128(_mm_cmpeq_epi8(a, vend), _mm_cmpeq_epi8(b, vend)));
                                        row[j >> 5] = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(a, pat)) | ((uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(b, pat)) << 16);
                                    }
                                    vend_seen = _mm_movemask_epi8(any_vend) != 0;
#endif
                                    for (; j < nb; ++j) {
                                        vend_seen |= gp[j] == -127;
                                        row[j >> 5] |= (uint32_t)(gp[j] == (int8_t)(4 + (j & 1))) << (j & 31);
                                    }
                                } else {
                                    for (uint32_t k = 0; k < S; ++k) {
                                        const size_t sp = co.sample_positions[k] * 2;
                                        if (sp + 2 > bytes) die("truncated BCF");
                                        const int8_t g0 = gp[sp], g1 = gp[sp + 1];
                                        vend_seen |= (g0 == -127) | (g1 == -127);
                                        row[k >> 4] |= ((uint32_t)(g0 == 4) | ((uint32_t)(g1 == 5) << 1)) << ((2 * k) & 31);
                                    }
                                }
                                if (vend_seen && problem.empty()) problem = "Inconsistent number of alleles";
                                d.p += bytes;
                                continue;
                            }
                            for (uint32_t k = 0; k < S; ++k) {
                                Cursor g{d.p + co.sample_positions[k] * 2 * es, d.p + bytes};
                                int32_t g0 = g.tint(vt), g1 = g.tint(vt);
                                // a vector shorter than the ploidy of the record is padded with END_OF_VECTOR (BCF2.2, 6.3.3); rust-htslib
                                // trims it, so genotype.len() < 2 and the assertion of haplotype.rs:32 fires
                                const int32_t vend = vt == 1 ? -127 : (vt == 2 ? -32767 : INT32_MIN + 1);
                                if ((g0 == vend || g1 == vend) && problem.empty()) problem = "Inconsistent number of alleles";
                                if (g0 == 4) row[(2 * k) >> 5] |= 1u << ((2 * k) & 31);          // Unphased(1), haplotype.rs:34-37
                                if (g1 == 5) row[(2 * k + 1) >> 5] |= 1u << ((2 * k + 1) & 31);  // Phased(1),   haplotype.rs:38-41
                            }
                        }
                        d.p += bytes;
                    }
                    if (!have_gt && S && problem.empty()) problem = "called `Result::unwrap()` on an `Err` value: missing GT";  // haplotype.rs:24
                }
            }
        } catch (const std::exception& e) {  // only under the test shim, where die() throws
            std::lock_guard<std::mutex> lk(err_mu);
            if (err.empty()) err = e.what();
        }
    };
    {
        const unsigned nt = std::max(1u, std::min<unsigned>(std::max(1u, o.threads), (unsigned)(pending.size() / 256 + 1)));
        if (nt == 1) decode();
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(decode);
            for (auto& t : th) t.join();
        }
        if (!err.empty()) die(err);
    }
    lap("pass 2 (GT -> carrier bits)");
    if (!std::is_sorted(co.records.begin(), co.records.end(), [](const Record& a, const Record& b) { return a.pos < b.pos; }))
        die("the BCF is not sorted by position (an indexed BCF always is)");
    return co;
}

}  // namespace
