// fasta.hpp -- FASTA windows through the .fai index (main.rs:156-161)
// Host side of find-tfbs-b200 (see driver.cpp for the map); header-only, one translation unit.
#pragma once
#include "options.hpp"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// FASTA through the .fai index (bio::io::fasta::IndexedReader)
// ---------------------------------------------------------------------------------------------------------------
struct Fasta {
    std::ifstream f;
    uint64_t len = 0, offset = 0, line_bases = 1, line_bytes = 1;
    std::string chrom_;
    Fasta(const std::string& path, const std::string& chrom) : f(path, std::ios::binary), chrom_(chrom) {
        if (!f) die("Error while opening the reference genome '" + path + "'");
        std::ifstream fai(path + ".fai");
        if (!fai) die("Error while opening the reference genome '" + path + "': missing .fai index");
        std::string line;
        bool found = false;
        while (std::getline(fai, line)) {
            auto x = split(line, '\t');
            if (x.size() >= 5 && x[0] == chrom) {
                len = strtoull(x[1].c_str(), nullptr, 10);
                offset = strtoull(x[2].c_str(), nullptr, 10);
                line_bases = std::max<uint64_t>(1, strtoull(x[3].c_str(), nullptr, 10));
                line_bytes = std::max<uint64_t>(1, strtoull(x[4].c_str(), nullptr, 10));
                found = true;
                break;
            }
        }
        if (!found) die("Error while seeking in reference genome file");
    }
    // [start, stop).  bio 0.28's IndexedReader::read refuses an interval that ends behind the contig ("FASTA read interval was out of
    // bounds") and the reference panics through .expect (main.rs:157-159): a peak within Lmax - 1 of the contig end kills its worker
    // there, and ends this program with the same message here.
    void fetch(uint64_t start, uint64_t stop, std::vector<uint8_t>* out) {
        if (stop > len || start > stop)
            die("Error while reading in reference genome file " + chrom_ + ":" + std::to_string(start) + "-" + std::to_string(stop - 1) +
                ": FASTA read interval was out of bounds");
        uint64_t pos = start;
        while (pos < stop) {
            uint64_t ln = pos / line_bases, col = pos % line_bases, take = std::min(stop - pos, line_bases - col);
            f.seekg((std::streamoff)(offset + ln * line_bytes + col));
            size_t old = out->size();
            out->resize(old + take);
            f.read((char*)out->data() + old, (std::streamsize)take);
            if ((uint64_t)f.gcount() != take) die("Error while reading in reference genome file");
            pos += take;
        }
    }
};

}  // namespace
