// driver.cpp -- find-tfbs-b200: the reference's command line on top of libtfbs_b200.so.
//
// Host side of the hot path, mirroring the reference (Helkafen/find-tfbs) so that a user can switch binaries:
//   options            src/main.rs:169-227   (long options are the API; the colliding short flags -n / -t are not offered)
//   PWM + thresholds   src/pattern.rs:13-117
//   BED + merge        src/bed.rs:9-60, src/range.rs:18-87
//   samples            src/main.rs:293-313
//   per region glue    src/main.rs:395-436   (halo :404-407, FASTA window :156-161, select_inner_peaks :62-72,
//                                              BCF fetch + GT rule src/haplotype.rs:13-62,78-79)
//   rows               src/main.rs:415-429, counts_as_genotypes :439-498 (second half: classes, dosage, COUNTS, freqs, maf)
//   writer             src/main.rs:264-290   (BGZF .part file, rename; --tabix runs tabix on it)
// Everything between "window + records in memory" and "(left, right) count vectors" runs on the GPU through the C ABI
// (include/tfbs.h).  There is no CPU implementation of that part in this program.
//
// Third-party formats the reference reads through crates (rust-htslib 0.26.1, bio 0.28.2, bgzip 0.0.3) are decoded here
// directly: BGZF (gzip members), BCF2.2, .fai-indexed FASTA, BED.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/tfbs.h"

namespace {

[[noreturn]] void die(const std::string& msg) {
#ifdef TFBS_DRIVER_TEST_SHIM
    throw std::runtime_error(msg);  // the CPU tests of the host logic catch the "panic"
#else
    fprintf(stderr, "find-tfbs-b200: %s\n", msg.c_str());
    exit(101);  // a Rust panic exits with 101
#endif
}

struct Range {
    uint64_t start, end;  // inclusive (range.rs:4-8)
    bool overlaps(const Range& o) const { return (o.start >= start && o.start <= end) || (o.end >= start && o.end <= end); }  // range.rs:18-21
    bool operator==(const Range& o) const { return start == o.start && end == o.end; }
};

// ---------------------------------------------------------------------------------------------------------------
// options
// ---------------------------------------------------------------------------------------------------------------
struct Options {
    std::string chromosome, bcf, output, reference, pwm_file, threshold_dir, samples_file;
    std::string audit_file;  // --audit: threshold ties, truncated and overwritten haplotypes (tfbs_audit_block), tab-separated
    std::vector<std::string> beds, pwm_names;
    float pwm_threshold = 0;
    bool forward_only = false, tabix = false, verbose = false, has_samples = false, plain_text = false;
    uint32_t min_maf = 0, threads = 1, chunk = 2000;
    uint64_t after_position = 0;
    std::vector<int> devices{0};
    bool use_index = true;   // --no_index: ignore <bcf>.csi and scan the whole BCF
};

std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    size_t p = 0;
    for (;;) {
        size_t q = s.find(sep, p);
        out.push_back(s.substr(p, q == std::string::npos ? std::string::npos : q - p));
        if (q == std::string::npos) break;
        p = q + 1;
    }
    return out;
}

void usage() {
    puts("find-tfbs-b200 1.0.1 (B200-native hot path)\n"
         "USAGE: find-tfbs-b200 --chromosome CHROM --input IN.bcf --output OUT.vcf.gz --reference REF.fa --bed A.bed[,B.bed]\n"
         "         --pwm_names NAME[,NAME] --pwm_file PWM.txt --pwm_threshold_directory DIR --pwm_threshold P\n"
         "         [--forward_only] [--threads N] [--min_maf N] [--after_position POS] [--samples FILE] [--tabix] [--verbose]\n"
         "         [--devices 0,1,...] [--chunk REGIONS_PER_BLOCK] [--plain] [--audit AUDIT.tsv] [--no_index]");
}

Options parse_args(int argc, char** argv) {
    Options o;
    std::map<std::string, std::string> kv;
    std::set<std::string> flags{"forward_only", "tabix", "verbose", "plain", "help", "no_index"};
    std::map<std::string, std::string> shorts{{"-c", "chromosome"}, {"-i", "input"}, {"-o", "output"}, {"-r", "reference"}, {"-b", "bed"},
                                              {"-p", "pwm_file"}, {"-f", "forward_only"}, {"-m", "min_maf"}, {"-s", "samples"},
                                              {"-z", "tabix"}, {"-v", "verbose"}};
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], key, val;
        bool has_val = false;
        if (a.rfind("--", 0) == 0) {
            size_t eq = a.find('=');
            key = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) { val = a.substr(eq + 1); has_val = true; }
        } else if (shorts.count(a)) key = shorts[a];
        else die("error: Found argument '" + a + "' which wasn't expected");
        if (flags.count(key)) { kv[key] = "1"; continue; }
        if (!has_val) {
            if (i + 1 >= argc) die("error: The argument '--" + key + "' requires a value");
            val = argv[++i];
        }
        kv[key] = val;
    }
    if (kv.count("help")) { usage(); exit(0); }
    auto req = [&](const char* k) -> std::string {
        if (!kv.count(k)) { usage(); die(std::string("error: The following required argument was not provided: --") + k); }
        return kv[k];
    };
    o.chromosome = req("chromosome");
    o.bcf = req("input");
    o.output = req("output");
    o.reference = req("reference");
    o.beds = split(req("bed"), ',');
    o.pwm_names = split(req("pwm_names"), ',');
    o.pwm_file = req("pwm_file");
    o.threshold_dir = req("pwm_threshold_directory");
    {
        char* e = nullptr;
        std::string t = req("pwm_threshold");
        o.pwm_threshold = strtof(t.c_str(), &e);
        if (e == t.c_str() || *e) die("Cannot parse MAF");  // sic, main.rs:195
    }
    o.forward_only = kv.count("forward_only");
    o.tabix = kv.count("tabix");
    o.verbose = kv.count("verbose");
    o.plain_text = kv.count("plain");
    o.use_index = !kv.count("no_index");
    auto num = [&](const char* k, uint64_t dflt, const char* what) -> uint64_t {
        if (!kv.count(k)) return dflt;
        char* e = nullptr;
        unsigned long long v = strtoull(kv[k].c_str(), &e, 10);
        if (e == kv[k].c_str() || *e) die(what);
        return v;
    };
    o.min_maf = (uint32_t)num("min_maf", 0, "Cannot parse MAF");
    o.threads = (uint32_t)num("threads", 1, "Cannot parse thread number");
    if (kv.count("threads") && o.threads < 1) die("Wrong number of threads");
    o.after_position = num("after_position", 0, "Cannot parse after_position");
    o.chunk = (uint32_t)std::max<uint64_t>(1, num("chunk", 2000, "Cannot parse chunk"));
    if (kv.count("samples")) { o.has_samples = true; o.samples_file = kv["samples"]; }
    if (kv.count("audit")) o.audit_file = kv["audit"];
    if (kv.count("devices")) {
        o.devices.clear();
        for (auto& d : split(kv["devices"], ',')) o.devices.push_back(atoi(d.c_str()));
    }
    return o;
}

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

// ---------------------------------------------------------------------------------------------------------------
// BED (bed.rs) + merge (range.rs)
// ---------------------------------------------------------------------------------------------------------------
std::vector<Range> load_bed(const std::string& path, const std::string& chrom) {
    std::ifstream f(path);
    if (!f) die("Bed file " + path + " does not exist");
    std::vector<Range> xs;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        auto fld = split(line, '\t');
        if (fld.size() < 3) die("malformed BED line in " + path + ": " + line);
        char *e1 = nullptr, *e2 = nullptr;
        uint64_t s = strtoull(fld[1].c_str(), &e1, 10), e = strtoull(fld[2].c_str(), &e2, 10);
        if (fld[1].empty() || fld[2].empty() || *e1 || *e2) die("malformed BED line in " + path + ": " + line);
        if (fld[0] == chrom) xs.push_back(Range{s, e});  // start/end used as an inclusive range (bed.rs:15)
    }
    return xs;
}

std::vector<Range> merge_ranges(std::vector<Range> raw) {  // range.rs:43-87
    std::stable_sort(raw.begin(), raw.end(), [](const Range& a, const Range& b) { return a.start < b.start; });
    std::vector<Range> out;
    for (const Range& r : raw) {
        if (!out.empty() && out.back().overlaps(r)) {
            out.back().start = std::min(out.back().start, r.start);
            out.back().end = std::max(out.back().end, r.end);
        } else out.push_back(r);
    }
    return out;
}

std::string basename_of(const std::string& s) {
    size_t p = s.find_last_of('/');
    return p == std::string::npos ? s : s.substr(p + 1);
}

// ---------------------------------------------------------------------------------------------------------------
// gzip / BGZF
// ---------------------------------------------------------------------------------------------------------------
// A file mapped read-only: pages are read when they are touched, so a reader that follows the index only pays for the blocks it uses.
class MappedFile {
public:
    MappedFile(const std::string& path, const char* what) {
        fd_ = open(path.c_str(), O_RDONLY);
        if (fd_ < 0) die(std::string(what) + " " + path);
        struct stat st;
        if (fstat(fd_, &st) != 0) die(std::string(what) + " " + path);
        size_ = (size_t)st.st_size;
        if (size_) {
            void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) die(std::string(what) + " " + path + " (mmap failed)");
            data_ = (const uint8_t*)m;
        }
    }
    ~MappedFile() {
        if (data_) munmap((void*)data_, size_);
        if (fd_ >= 0) close(fd_);
    }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    const uint8_t* data() const { return data_; }
    size_t size() const { return size_; }

private:
    int fd_ = -1;
    const uint8_t* data_ = nullptr;
    size_t size_ = 0;
};

// Inflates gzip members one after the other; stops early once `limit` bytes are there (the BCF header is read this way).
std::vector<uint8_t> gunzip_members(const uint8_t* in, size_t n_in, const std::string& what, size_t limit = SIZE_MAX) {
    std::vector<uint8_t> out;
    size_t off = 0;
    while (off < n_in && out.size() < limit) {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 16) != Z_OK) die("zlib initialisation failed");
        zs.next_in = const_cast<Bytef*>(in + off);
        zs.avail_in = (uInt)std::min<size_t>(n_in - off, 1u << 30);
        int rc;
        do {
            size_t old = out.size();
            out.resize(old + (1u << 17));
            zs.next_out = out.data() + old;
            zs.avail_out = 1u << 17;
            rc = inflate(&zs, Z_NO_FLUSH);
            out.resize(old + ((1u << 17) - zs.avail_out));
            if (rc != Z_OK && rc != Z_STREAM_END) die("corrupt compressed stream in " + what);
        } while (rc != Z_STREAM_END);
        off += zs.total_in;
        inflateEnd(&zs);
    }
    return out;
}

// BGZF: every gzip member carries its own size (BSIZE in the 'BC' extra subfield) and its uncompressed size (ISIZE), so the
// members can be located without inflating and inflated independently on several threads.  Falls back to the serial reader for a
// plain gzip stream.
// Size of the BGZF member at in[off..): 0 if it is not one (plain gzip, truncated).
size_t bgzf_member_size(const uint8_t* in, size_t n_in, size_t off) {
    if (n_in - off < 28 || in[off] != 0x1f || in[off + 1] != 0x8b || !(in[off + 3] & 4)) return 0;
    const size_t xlen = in[off + 10] | (in[off + 11] << 8);
    size_t p = off + 12, bsize = 0;
    const size_t xend = p + xlen;
    if (xend > n_in) return 0;
    while (p + 4 <= xend) {
        const size_t slen = in[p + 2] | (in[p + 3] << 8);
        if (in[p] == 'B' && in[p + 1] == 'C' && slen == 2 && p + 6 <= xend) bsize = (size_t)(in[p + 4] | (in[p + 5] << 8)) + 1;
        p += 4 + slen;
    }
    if (bsize < 26 || off + bsize > n_in) return 0;
    return bsize;
}

std::vector<uint8_t> gunzip_bgzf(const uint8_t* in, size_t n_in, const std::string& what, unsigned threads) {
    struct Member { size_t off, csize, uoff; uint32_t isize; };
    std::vector<Member> ms;
    size_t off = 0, total = 0;
    while (off < n_in) {
        const size_t bsize = bgzf_member_size(in, n_in, off);
        if (!bsize) return gunzip_members(in, n_in, what);
        uint32_t isize;
        memcpy(&isize, in + off + bsize - 4, 4);
        ms.push_back(Member{off, bsize, total, isize});
        total += isize;
        off += bsize;
    }
    std::vector<uint8_t> out(total);
    std::atomic<size_t> next{0};
    std::atomic<bool> bad{false};
    auto work = [&] {
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= ms.size()) return;
            const Member& m = ms[k];
            if (m.isize == 0) continue;
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, 15 + 16) != Z_OK) { bad = true; return; }
            zs.next_in = const_cast<Bytef*>(in + m.off);
            zs.avail_in = (uInt)m.csize;
            zs.next_out = out.data() + m.uoff;
            zs.avail_out = m.isize;
            int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.total_out != m.isize) bad = true;
            inflateEnd(&zs);
        }
    };
    const unsigned nt = std::max(1u, std::min<unsigned>(threads, (unsigned)ms.size()));
    if (nt == 1) work();
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    if (bad) die("corrupt compressed stream in " + what);
    return out;
}

// BGZF writer: independent gzip members of <= 64 KiB with the BC extra field, terminated by the empty EOF block.
class BgzfWriter {
public:
    BgzfWriter(const std::string& path, unsigned threads) : f_(path, std::ios::binary), threads_(std::max(1u, threads)) {
        if (!f_) die("Could not create output file");
    }
    void write(const std::string& s) {
        buf_ += s;
        if (buf_.size() >= kBlock * 8 * threads_) flush(false);
    }
    void finish() {
        flush(true);
        const std::string eof = compress(nullptr, 0);  // EOF marker
        f_.write(eof.data(), (std::streamsize)eof.size());
        f_.close();
    }

private:
    static constexpr size_t kBlock = 0xff00;
    // the blocks are independent gzip members: compressed on several threads, written in order
    void flush(bool all) {
        const size_t n_full = buf_.size() / kBlock, n = n_full + ((all && buf_.size() % kBlock) ? 1 : 0);
        if (n == 0) return;
        std::vector<std::string> out(n);
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (;;) {
                size_t k = next.fetch_add(1);
                if (k >= n) return;
                out[k] = compress(buf_.data() + k * kBlock, std::min(kBlock, buf_.size() - k * kBlock));
            }
        };
        const unsigned nt = (unsigned)std::min<size_t>(threads_, n);
        if (nt <= 1) work();
        else {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
            for (auto& t : th) t.join();
        }
        for (const std::string& blk : out) f_.write(blk.data(), (std::streamsize)blk.size());
        buf_.erase(0, std::min(buf_.size(), n * kBlock));
    }
    static std::string compress(const char* data, size_t n) {
        std::string out(0x10000 + 64, '\0');
        uint8_t* o = (uint8_t*)&out[0];
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = (Bytef*)data;
        zs.avail_in = (uInt)n;
        zs.next_out = o + 18;
        zs.avail_out = (uInt)(out.size() - 18 - 8);
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) die("deflate failed");
        size_t clen = zs.total_out;
        deflateEnd(&zs);
        const uint8_t hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
        memcpy(o, hdr, 12);
        o[12] = 'B'; o[13] = 'C'; o[14] = 2; o[15] = 0;
        size_t bsize = clen + 25;  // total block size - 1
        o[16] = bsize & 0xff; o[17] = (bsize >> 8) & 0xff;
        uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)data, (uInt)n);
        uint32_t isize = (uint32_t)n;
        memcpy(o + 18 + clen, &crc, 4);
        memcpy(o + 22 + clen, &isize, 4);
        out.resize(clen + 26);
        return out;
    }
    std::ofstream f_;
    std::string buf_;
    unsigned threads_;
};

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

void check_letters(const std::string& s) {  // util.rs:4-16
    for (unsigned char l : s)
        if (!(l == 65 || l == 97 || l == 67 || l == 99 || l == 71 || l == 103 || l == 84 || l == 116 || l == 78 || l == 110))
            die("Unknown nucleotide " + std::to_string((int)l));
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
    std::vector<uint8_t> d = gunzip_bgzf(f.data(), f.size(), bcf_path + ".csi", 1);
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
    std::vector<uint8_t> raw;
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
    if (first_record > raw.size()) die("truncated BCF");
    Cursor c{raw.data() + first_record, raw.data() + raw.size()};
    // pass 1 (serial, a few bytes per record): positions, alleles, carrier rows; the genotype blocks are only located
    struct Pending { uint32_t row; const uint8_t* indiv; uint32_t l_indiv, n_fmt, n_sample; };
    std::vector<Pending> pending;
    while (c.p < c.e) {
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
        if (r.n_allele < 2) die("index out of bounds: the len is " + std::to_string(r.n_allele) + " but the index is 1");  // haplotype.rs:22
        r.ref = s.tstr();
        r.alt = s.tstr();
        check_letters(r.ref);  // haplotype.rs:21-22 convert alleles[0] and [1] of every record
        check_letters(r.alt);
        r.carrier_row = UINT32_MAX;
        if (r.n_allele == 2) {
            r.carrier_row = (uint32_t)pending.size();
            pending.push_back(Pending{r.carrier_row, indiv, l_indiv, n_fmt, n_sample});
        } else {
            printf("Unusual number of alleles: %u\n", r.n_allele);  // haplotype.rs:53-55
        }
        co.max_rlen = std::max(co.max_rlen, std::max(1, r.rlen));
        co.records.push_back(std::move(r));
    }
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
                    for (uint32_t f = 0; f < pd.n_fmt; ++f) {
                        int kt, vt; uint32_t kl, vl;
                        d.desc(&kt, &kl);
                        int32_t key = d.tint(kt);
                        d.desc(&vt, &vl);
                        size_t bytes = Cursor::tsize(vt) * vl * (size_t)pd.n_sample;
                        d.need(bytes);
                        if (key == gt_key && vt >= 1 && vt <= 3) {
                            have_gt = true;
                            if (vl != 2 && S) die("Inconsistent number of alleles");  // haplotype.rs:32
                            const size_t es = Cursor::tsize(vt);
                            for (uint32_t k = 0; k < S; ++k) {
                                Cursor g{d.p + co.sample_positions[k] * 2 * es, d.p + bytes};
                                int32_t g0 = g.tint(vt), g1 = g.tint(vt);
                                if (g0 == 4) row[(2 * k) >> 5] |= 1u << ((2 * k) & 31);          // Unphased(1), haplotype.rs:34-37
                                if (g1 == 5) row[(2 * k + 1) >> 5] |= 1u << ((2 * k + 1) & 31);  // Phased(1),   haplotype.rs:38-41
                            }
                        }
                        d.p += bytes;
                    }
                    if (!have_gt && S) die("called `Result::unwrap()` on an `Err` value: missing GT");  // haplotype.rs:24
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
    if (!std::is_sorted(co.records.begin(), co.records.end(), [](const Record& a, const Record& b) { return a.pos < b.pos; }))
        die("the BCF is not sorted by position (an indexed BCF always is)");
    return co;
}

// ---------------------------------------------------------------------------------------------------------------
// FASTA through the .fai index (bio::io::fasta::IndexedReader)
// ---------------------------------------------------------------------------------------------------------------
struct Fasta {
    std::ifstream f;
    uint64_t len = 0, offset = 0, line_bases = 1, line_bytes = 1;
    Fasta(const std::string& path, const std::string& chrom) : f(path, std::ios::binary) {
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
    void fetch(uint64_t start, uint64_t stop, std::vector<uint8_t>* out) {  // [start, stop), clipped at the contig end
        stop = std::min(stop, len);
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
        b.carriers = co.carriers.data();
        b.n_carrier_rows = (uint32_t)(co.carriers.size() / co.pitch);
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
            v.carrier_row = it->carrier_row;
            bd->variants.push_back(v);
        }
        bd->n_records.push_back(nrec);
        bd->var_off.push_back((uint32_t)bd->variants.size());
    }
}

// ---------------------------------------------------------------------------------------------------------------
// rows: second half of counts_as_genotypes (main.rs:459-498) and the row text (main.rs:415-425)
// ---------------------------------------------------------------------------------------------------------------
struct RowText {
    bool keep = false;
    std::string info, genotypes;
};

template <class T>
RowText finalise_row(const T* l, const T* r, uint32_t S, uint32_t lowest, uint32_t highest, uint32_t min_maf) {
    RowText t;
    if (lowest == highest) return t;  // main.rs:456-458 (the library already filtered these)
    const uint32_t i1 = (lowest * 1000u * 3u + highest * 1000u) / 4u;  // :461
    const uint32_t i3 = (lowest * 1000u + highest * 1000u * 3u) / 4u;  // :462
    std::vector<uint32_t> all{lowest, highest};
    uint32_t zero = 0, one = 0, two = 0;
    const float lowest_f = (float)lowest, spread = (float)highest - lowest_f;
    t.genotypes.reserve((size_t)S * 12);
    char buf[48];
    for (uint32_t s = 0; s < S; ++s) {
        uint32_t x = (uint32_t)l[s] + (uint32_t)r[s];
        if (x == lowest) { t.genotypes += "\t0|0:0.0"; ++zero; }
        else if (x == highest) { t.genotypes += "\t1|1:2.0"; ++two; }
        else {
            if (std::find(all.begin(), all.end(), x) == all.end()) all.push_back(x);
            uint32_t x1000 = x * 1000u;
            if (x1000 < i1) { t.genotypes += "\t0|0"; ++zero; }
            else if (x1000 < i3) { t.genotypes += "\t0|1"; ++one; }
            else { t.genotypes += "\t1|1"; ++two; }
            volatile float num = ((float)x - lowest_f) * 2.0f;  // f32 steps as in :478
            float dosage = num / spread;
            snprintf(buf, sizeof buf, ":%.4f", (double)dosage);  // {:.4}
            t.genotypes += buf;
        }
    }
    uint32_t maf = (zero >= one && zero >= two) ? one + two : (two >= zero && two >= one) ? zero + one : zero + two;  // :482-489
    if (maf < min_maf) return t;  // main.rs:421
    std::sort(all.begin(), all.end());
    t.info = "COUNTS=";
    for (size_t i = 0; i < all.size(); ++i) t.info += (i ? "," : "") + std::to_string(all[i]);
    t.info += ";freqs=" + std::to_string(zero) + "/" + std::to_string(one) + "/" + std::to_string(two);
    t.keep = true;
    return t;
}

#ifndef TFBS_DRIVER_NO_MAIN
#define TF(call)                                                         \
    do {                                                                 \
        int rc_ = (call);                                                \
        if (rc_ != TFBS_OK) die(tfbs_last_error(ctx));                   \
    } while (0)

#endif

}  // namespace

#ifdef TFBS_DRIVER_TEST_SHIM
// Flat entry points over the host logic above, for the CPU tests (tests/test_driver_host.py).  Not part of the program.
extern "C" {
static thread_local std::string g_shim_err;
const char* drv_last_error() { return g_shim_err.c_str(); }

int drv_parse_weight(const char* s, int32_t* out) {
    try { *out = parse_weight(s); return 0; } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
}

int drv_parse_threshold_file(const char* path, float thr, int32_t* out) {
    try { return parse_threshold_file(path, thr, out) ? 1 : 0; } catch (const std::exception& e) { g_shim_err = e.what(); return -1; }
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
#endif

#ifndef TFBS_DRIVER_NO_MAIN
int main(int argc, char** argv) {
    Options o = parse_args(argc, argv);
    if (o.tabix && system("command -v tabix > /dev/null 2>&1") != 0) die("tabix cannot in found in PATH");  // main.rs:220-223
    auto t_start = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };

    std::vector<Pwm> pwms = parse_pwm_files(o);
    if (pwms.empty()) die("assertion failed: pwm_list.len() > 0");  // main.rs:238
    std::map<uint16_t, std::string> pwm_name;
    uint32_t largest = 0;
    for (const Pwm& p : pwms) {
        printf("PWM %s %d %s %zu\n", p.name.c_str(), p.min_score, p.direction == TFBS_DIR_P ? "P" : "N", p.w.size() / 4);
        pwm_name[p.pattern_id] = p.name;
        largest = std::max<uint32_t>(largest, (uint32_t)(p.w.size() / 4));
    }

    // load_peak_files (bed.rs:25-60)
    std::vector<std::vector<Range>> peak_map;
    std::vector<std::string> bed_names;
    std::vector<Range> all;
    for (const std::string& b : o.beds) {
        std::vector<Range> peaks = load_bed(b, o.chromosome), kept;
        uint64_t cover = 0;
        for (const Range& p : peaks) cover += p.end - p.start;
        printf("Loaded %s:\t %zu peaks covering %llu bp\n", b.c_str(), peaks.size(), (unsigned long long)cover);
        for (const Range& p : peaks)
            if (p.start >= o.after_position) kept.push_back(p);
        all.insert(all.end(), kept.begin(), kept.end());
        peak_map.push_back(kept);
        bed_names.push_back(basename_of(b));
    }
    std::vector<Range> merged = merge_ranges(all);
    printf("Merged all region files: %zu merged regions\n", merged.size());

    // one context per device; chunks of merged regions are dealt round-robin (regions are independent, main.rs:395-429)
    std::vector<tfbs_pattern> cpat(pwms.size());
    for (size_t i = 0; i < pwms.size(); ++i) {
        cpat[i].weights = pwms[i].w.data();
        cpat[i].len = (uint32_t)(pwms[i].w.size() / 4);
        cpat[i].min_score = pwms[i].min_score;
        cpat[i].pattern_id = pwms[i].pattern_id;
        cpat[i].direction = pwms[i].direction;
        cpat[i].kind = TFBS_PATTERN_PWM;
    }
    // the CUDA contexts are created (and the pattern tables compiled and uploaded) while the BCF is read
    std::atomic<uint64_t> us_create{0};
    auto now_us = [] { return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    if (o.devices.empty()) o.devices.push_back(0);
    std::vector<tfbs_ctx*> ctxs(o.devices.size(), nullptr);
    std::vector<std::thread> pre;
    for (size_t k = 0; k < o.devices.size(); ++k)
        pre.emplace_back([&, k] {
            const uint64_t tc = now_us();
            tfbs_ctx* ctx = nullptr;
            if (tfbs_create(o.devices[k], &ctx) != TFBS_OK) die(std::string(tfbs_last_error(nullptr)));
            TF(tfbs_set_patterns(ctx, cpat.data(), (uint32_t)cpat.size()));
            TF(tfbs_set_option(ctx, "rows_width", 0));
            ctxs[k] = ctx;
            us_create += now_us() - tc;
        });
    const double t_before_bcf = since();
    Cohort co = load_bcf(o);
    const double t_after_bcf = since();
    for (auto& t : pre) t.join();
    const uint32_t S = (uint32_t)co.samples.size();

    std::string chr = o.chromosome;  // main.rs:402
    for (size_t p; (p = chr.find("chr")) != std::string::npos;) chr.erase(p, 3);

    const std::string part = o.output + ".part";
    BgzfWriter* gz = o.plain_text ? nullptr : new BgzfWriter(part, o.threads);
    std::ofstream plain;
    if (o.plain_text) { plain.open(part, std::ios::binary); if (!plain) die("Could not create output file"); }
    auto emit = [&](const std::string& s) { if (gz) gz->write(s); else plain << s; };
    {
        std::string hdr = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";  // main.rs:320-324
        for (auto& s : co.samples) hdr += "\t" + s;
        emit(hdr + "\n");
    }

    const size_t n_chunks = (merged.size() + o.chunk - 1) / o.chunk;
    struct ChunkOut { std::vector<std::string> rows; std::string audit; };
    std::vector<ChunkOut> outs(n_chunks);
    std::atomic<size_t> next{0};
    std::atomic<uint64_t> total_cells{0}, total_hits{0};
    std::atomic<uint64_t> us_wait{0}, us_gpu{0}, us_sort{0}, us_format{0};
    auto worker = [&](size_t slot) {
        tfbs_ctx* ctx = ctxs[slot];
        // a builder thread prepares the next blocks (FASTA windows, inner regions, records) while the GPU works on the current one;
        // private readers per worker, like main.rs:345-346
        struct Ready { size_t c; std::unique_ptr<BlockData> bd; };
        std::mutex mu;
        std::condition_variable cv;
        std::deque<Ready> ready;
        bool finished = false;
        std::thread builder([&] {
            Fasta fa(o.reference, o.chromosome);
            for (;;) {
                size_t c = next.fetch_add(1);
                if (c >= n_chunks) break;
                std::unique_ptr<BlockData> bd(new BlockData());
                build_block(merged, c * o.chunk, std::min(merged.size(), (c + 1) * (size_t)o.chunk), peak_map, co, fa, largest, bd.get());
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return ready.size() < 2; });
                ready.push_back(Ready{c, std::move(bd)});
                cv.notify_all();
            }
            std::lock_guard<std::mutex> lk(mu);
            finished = true;
            cv.notify_all();
        });
        for (;;) {
            Ready item;
            uint64_t tw = now_us();
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !ready.empty() || finished; });
                if (ready.empty()) break;
                item = std::move(ready.front());
                ready.pop_front();
                cv.notify_all();
            }
            const size_t c = item.c, m0 = c * o.chunk, m1 = std::min(merged.size(), m0 + o.chunk);
            BlockData& bd = *item.bd;
            us_wait += now_us() - tw;
            uint64_t tg = now_us();
            tfbs_block blk = bd.view(co);
            if (o.audit_file.empty()) {
                TF(tfbs_submit_block(ctx, &blk));
            } else {
                // the audit scores the block twice (thresholds lowered by one, then as given) and leaves the rows of the normal run
                TF(tfbs_upload_block(ctx, &blk));
                tfbs_audit au;
                TF(tfbs_audit_block(ctx, &au));
                if (au.truncated) die("--audit: the match buffer overflowed; use a smaller --chunk");
                std::string& out = outs[c].audit;
                const uint32_t H = 2 * au.n_samples;
                auto region_name = [&](uint32_t r) { return std::to_string(merged[m0 + r].start) + "-" + std::to_string(merged[m0 + r].end); };
                for (uint64_t i = 0; i < au.n_ties; ++i) {
                    const uint32_t r = au.tie_region[i], g = au.tie_group[i];
                    uint32_t first = UINT32_MAX, members = 0;  // the group's first haplotype names it
                    for (uint32_t h = 0; h < H; ++h)
                        if (au.hap_group[(size_t)r * H + h] == g) { if (first == UINT32_MAX) first = h; ++members; }
                    const Pwm& pw = pwms[au.tie_pattern_index[i]];
                    out += "tie\t" + chr + "\t" + region_name(r) + "\t" + pw.name + "\t" + (pw.direction == TFBS_DIR_P ? "P" : "N") + "\t" +
                           std::to_string(au.tie_start[i]) + "\t" + std::to_string(pw.min_score) + "\t" +
                           (g == 0 ? std::string("reference") : co.samples[first / 2] + (first % 2 ? ":R" : ":L")) + "\t" + std::to_string(members) + "\n";
                }
                for (uint32_t r = 0; r < au.n_regions; ++r)
                    for (uint32_t h = 0; h < H; ++h) {
                        const uint8_t fl = au.hap_flags[(size_t)r * H + h];
                        if (fl & TFBS_HAP_TRUNCATED)
                            out += "truncated\t" + chr + "\t" + region_name(r) + "\t" + co.samples[h / 2] + (h % 2 ? ":R" : ":L") + "\n";
                        if (fl & TFBS_HAP_OVERWRITTEN)
                            out += "overwritten\t" + chr + "\t" + region_name(r) + "\t" + co.samples[h / 2] + (h % 2 ? ":R" : ":L") + "\n";
                    }
            }
            tfbs_rows rows;
            TF(tfbs_collect(ctx, &rows));
            us_gpu += now_us() - tg;
            uint64_t ts = now_us();
            tfbs_stats st;
            tfbs_get_stats(ctx, &st);
            total_cells += st.nominal_cells;
            total_hits += st.n_hits;
            // canonical order inside a region: (bed file, inner range, pattern_id); the reference's order is HashMap::drain()
            std::vector<uint64_t> order(rows.n_rows);
            for (uint64_t i = 0; i < rows.n_rows; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
                if (rows.region[a] != rows.region[b]) return rows.region[a] < rows.region[b];
                const tfbs_inner_region& ia = bd.inner[rows.inner[a]];
                const tfbs_inner_region& ib = bd.inner[rows.inner[b]];
                if (ia.bed_index != ib.bed_index) return ia.bed_index < ib.bed_index;
                if (ia.start != ib.start) return ia.start < ib.start;
                if (ia.end != ib.end) return ia.end < ib.end;
                return rows.pattern_id[a] < rows.pattern_id[b];
            });
            us_sort += now_us() - ts;
            uint64_t tf = now_us();
            // row text on --threads host threads (the reference formats inside its worker threads, main.rs:415-425)
            std::vector<std::string> text(order.size());
            auto format_range = [&](size_t a, size_t b) {
                for (size_t k = a; k < b; ++k) {
                    const uint64_t i = order[k];
                    // counts arrive in the narrowest type that holds them (option rows_width = 0, tfbs_rows.count_bytes)
                    RowText t = rows.count_bytes == 1 ? finalise_row((const uint8_t*)rows.left + i * S, (const uint8_t*)rows.right + i * S, S, rows.vmin[i], rows.vmax[i], o.min_maf)
                              : rows.count_bytes == 2 ? finalise_row((const uint16_t*)rows.left + i * S, (const uint16_t*)rows.right + i * S, S, rows.vmin[i], rows.vmax[i], o.min_maf)
                                                      : finalise_row(rows.left + i * S, rows.right + i * S, S, rows.vmin[i], rows.vmax[i], o.min_maf);
                    if (!t.keep) continue;
                    const tfbs_inner_region& ir = bd.inner[rows.inner[i]];
                    // POS is filled in by the writer (a running counter, main.rs:329,424-425)
                    text[k] = "\t" + bed_names[ir.bed_index] + "," + pwm_name.at(rows.pattern_id[i]) + "," + std::to_string(ir.start) + "-" +
                              std::to_string(ir.end) + "\t.\t.\t.\tPASS\t" + t.info + "\tGT:DS" + t.genotypes + "\n";
                }
            };
            const size_t nt = std::max<size_t>(1, std::min<size_t>(o.threads, order.size() / 64 + 1));
            if (nt == 1) format_range(0, order.size());
            else {
                std::vector<std::thread> ft;
                for (size_t t = 0; t < nt; ++t) ft.emplace_back(format_range, order.size() * t / nt, order.size() * (t + 1) / nt);
                for (auto& t : ft) t.join();
            }
            for (std::string& row : text)
                if (!row.empty()) outs[c].rows.push_back(std::move(row));
            us_format += now_us() - tf;
            if (o.verbose)
                printf("\nChunk %zu/%zu\tregions %zu-%zu\t%llu haplotypes\t%llu hits\n", c + 1, n_chunks, m0, m1, (unsigned long long)st.n_groups,
                       (unsigned long long)st.n_hits);
        }
        builder.join();
        tfbs_destroy(ctx);
    };
    if (o.devices.size() <= 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (size_t k = 0; k < o.devices.size(); ++k) th.emplace_back(worker, k);
        for (auto& t : th) t.join();
    }
    const double t_after_gpu = since();
    uint64_t fake_position = 1;
    for (const ChunkOut& co2 : outs)
        for (const std::string& row : co2.rows) emit(chr + "\t" + std::to_string(fake_position++) + row);
    if (gz) { gz->finish(); delete gz; } else plain.close();
    if (!o.audit_file.empty()) {
        // tie: a window scoring exactly min_score (not a hit, pattern.rs:151); truncated: haplotype.rs:144-149; overwritten: the
        // haplotype's entry of the sequence-keyed map was replaced (haplotype.rs:84), it is counted with the reference haplotype and
        // the reference program's choice between the colliding entries depends on HashMap order
        std::ofstream af(o.audit_file, std::ios::binary);
        if (!af) die("Could not create audit file");
        af << "#tie\tCHROM\tREGION\tPWM\tSTRAND\tSTART\tMIN_SCORE\tGROUP\tHAPLOTYPES\n#truncated|overwritten\tCHROM\tREGION\tHAPLOTYPE\n";
        for (const ChunkOut& co2 : outs) af << co2.audit;
    }
    if (rename(part.c_str(), o.output.c_str()) != 0) die("Could not rename " + part + " into " + o.output);
    if (o.tabix) {
        std::string cmd = "tabix -f -p vcf '" + o.output + "'";
        if (system(cmd.c_str()) == 0) printf("Tabixed file %s\n", o.output.c_str());
        else printf("Failed to tabix file %s\n", o.output.c_str());
    }
    double secs = since();
    printf("phases: PWM+BED %.2f s, BCF %.2f s, blocks+GPU+rows %.2f s (context %.2f, waiting for blocks %.2f, submit+collect %.2f, sort %.2f, "
           "row text %.2f; summed over devices), VCF write %.2f s\n", t_before_bcf, t_after_bcf - t_before_bcf, t_after_gpu - t_after_bcf,
           us_create / 1e6, us_wait / 1e6, us_gpu / 1e6, us_sort / 1e6, us_format / 1e6, secs - t_after_gpu);
    printf("%zu merged regions, %llu rows, %llu hits, %.3e nominal cells in %.2f s\nEnd of program.\n", merged.size(),
           (unsigned long long)(fake_position - 1), (unsigned long long)total_hits.load(), (double)total_cells.load(), secs);
    return 0;
}
#endif
