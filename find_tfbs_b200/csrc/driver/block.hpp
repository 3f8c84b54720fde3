// block.hpp -- what process_peak gathers for a chunk of merged regions (main.rs:395-413) as a tfbs_block
// Host side of find-tfbs-b200 (see driver.cpp for the map); header-only, one translation unit.
#pragma once
#include "bcf.hpp"
#include "fasta.hpp"
#include "bed.hpp"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// block building: what process_peak gathers for a merged region (main.rs:395-413)
// ---------------------------------------------------------------------------------------------------------------
struct BlockData {
    std::vector<int64_t> region_start, region_end;
    std::vector<uint64_t> ref_off{0};
    std::vector<uint8_t> ref_bases;
    std::vector<uint32_t> inner_off{0};
    std::vector<tfbs_inner_region> inner;
    std::vector<uint32_t> var_off{0};
    std::vector<tfbs_variant> variants;
    std::vector<uint8_t> alleles;
    std::vector<uint32_t> n_records;  // all records of the window, incl. non-biallelic ("variants" of main.rs:435)
    std::vector<uint32_t> carriers;   // the carrier rows of this block's records only ([rows][pitch]): a block does not ship the cohort's matrix
    std::unordered_map<uint32_t, uint32_t> row_of;  // cohort row -> row of this block (windows of neighbouring regions share records)
    tfbs_block view(const Cohort& co) const {
        tfbs_block b;
        memset(&b, 0, sizeof b);
        b.n_regions = (uint32_t)region_start.size();
        b.n_samples = (uint32_t)co.samples.size();
        b.region_start = region_start.data();
        b.region_end = region_end.data();
        b.ref_off = ref_off.data();
        b.ref_bases = ref_bases.data();
        b.inner_off = inner_off.data();
        b.inner = inner.data();
        b.var_off = var_off.data();
        b.variants = variants.data();
        b.allele_bases = alleles.data();
        b.allele_bytes = alleles.size();
        b.carriers = carriers.data();
        b.n_carrier_rows = (uint32_t)(carriers.size() / co.pitch);
        b.carrier_pitch = co.pitch;
        return b;
    }
};

void build_block(const std::vector<Range>& merged, size_t m0, size_t m1, const std::vector<std::vector<Range>>& peak_map, const Cohort& co,
                 Fasta& fa, uint32_t largest, BlockData* bd) {
    for (size_t m = m0; m < m1; ++m) {
        const Range& mr = merged[m];
        if (mr.start + 1 < largest) die("attempt to subtract with overflow");  // main.rs:407 in a debug build
        Range ext{mr.start - largest + 1, mr.end + largest - 1};
        bd->region_start.push_back((int64_t)ext.start);
        bd->region_end.push_back((int64_t)ext.end);
        fa.fetch(ext.start, ext.end + 1, &bd->ref_bases);  // main.rs:157
        bd->ref_off.push_back(bd->ref_bases.size());
        // select_inner_peaks (main.rs:62-72): p.overlaps(merged) -- asymmetric; equal ranges of one file collapse into a multiplicity
        for (uint32_t b = 0; b < peak_map.size(); ++b) {
            size_t first = bd->inner.size();
            for (const Range& p : peak_map[b]) {
                if (!p.overlaps(mr)) continue;
                bool dup = false;
                for (size_t k = first; k < bd->inner.size(); ++k)
                    if ((uint64_t)bd->inner[k].start == p.start && (uint64_t)bd->inner[k].end == p.end) { bd->inner[k].multiplicity++; dup = true; break; }
                if (!dup) bd->inner.push_back(tfbs_inner_region{(int64_t)p.start, (int64_t)p.end, b, 1});
            }
        }
        bd->inner_off.push_back((uint32_t)bd->inner.size());
        // reader.fetch(rid, start, end + 1) (haplotype.rs:79): records overlapping [start, end + 1)
        int64_t ws = (int64_t)ext.start, we = (int64_t)ext.end;
        auto lo = std::lower_bound(co.records.begin(), co.records.end(), ws - co.max_rlen, [](const Record& r, int64_t v) { return r.pos < v; });
        uint32_t nrec = 0;
        for (auto it = lo; it != co.records.end() && it->pos <= we; ++it) {
            if (it->pos + std::max(1, it->rlen) <= ws) continue;
            ++nrec;
            if (!it->problem.empty()) die(it->problem);  // the reference panics on such a record once a region fetches it (haplotype.rs:21-32)
            if (it->n_allele != 2) printf("Unusual number of alleles: %u\n", it->n_allele);  // haplotype.rs:53-55
            if (it->carrier_row == UINT32_MAX) continue;  // not biallelic: counted, not used (haplotype.rs:27,53-55)
            tfbs_variant v;
            memset(&v, 0, sizeof v);
            v.pos = it->pos;
            v.ref_off = (uint32_t)bd->alleles.size();
            v.ref_len = (uint32_t)it->ref.size();
            bd->alleles.insert(bd->alleles.end(), it->ref.begin(), it->ref.end());
            v.alt_off = (uint32_t)bd->alleles.size();
            v.alt_len = (uint32_t)it->alt.size();
            bd->alleles.insert(bd->alleles.end(), it->alt.begin(), it->alt.end());
            auto found = bd->row_of.find(it->carrier_row);
            if (found == bd->row_of.end()) {
                found = bd->row_of.emplace(it->carrier_row, (uint32_t)(bd->carriers.size() / co.pitch)).first;
                const uint32_t* src = co.carriers.data() + (size_t)it->carrier_row * co.pitch;
                bd->carriers.insert(bd->carriers.end(), src, src + co.pitch);
            }
            v.carrier_row = found->second;
            bd->variants.push_back(v);
        }
        bd->n_records.push_back(nrec);
        bd->var_off.push_back((uint32_t)bd->variants.size());
    }
}

}  // namespace
