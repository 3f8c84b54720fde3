// bgzf.hpp -- memory-mapped files, gzip / BGZF reader (parallel inflate) and BGZF writer
// Host side of find-tfbs-b200 (see driver.cpp for the map); header-only, one translation unit.
#pragma once
#include "common.hpp"

namespace {

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

// Bytes that are not zeroed when the vector is sized: the inflated BCF of a cohort is gigabytes that every inflate thread
// overwrites anyway (zero-filling them first costs a serial pass over the whole buffer).
template <class T>
struct NoInit {
    using value_type = T;
    NoInit() = default;
    template <class U> NoInit(const NoInit<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(::operator new(n * sizeof(T))); }
    void deallocate(T* p, size_t) { ::operator delete(p); }
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U;
        else ::new ((void*)p) U(std::forward<A>(a)...);
    }
    template <class U> bool operator==(const NoInit<U>&) const { return true; }
    template <class U> bool operator!=(const NoInit<U>&) const { return false; }
};
using RawBytes = std::vector<uint8_t, NoInit<uint8_t>>;

RawBytes gunzip_bgzf(const uint8_t* in, size_t n_in, const std::string& what, unsigned threads) {
    struct Member { size_t off, csize, uoff; uint32_t isize; };
    std::vector<Member> ms;
    size_t off = 0, total = 0;
    while (off < n_in) {
        const size_t bsize = bgzf_member_size(in, n_in, off);
        if (!bsize) { const std::vector<uint8_t> v = gunzip_members(in, n_in, what); return RawBytes(v.begin(), v.end()); }
        uint32_t isize;
        memcpy(&isize, in + off + bsize - 4, 4);
        ms.push_back(Member{off, bsize, total, isize});
        total += isize;
        off += bsize;
    }
    RawBytes out(total);
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
    BgzfWriter(const std::string& path, unsigned threads, int level = 6) : f_(path, std::ios::binary), threads_(std::max(1u, threads)), level_(level) {
        if (!f_) die("Could not create output file");
    }
    void write(const std::string& s) { write(s.data(), s.size()); }
    void write(const char* p, size_t n) {  // pieces of a line go straight into the block buffer
        buf_.append(p, n);
        if (buf_.size() >= kBlock * 8 * threads_) flush(false);
    }
    // text -> complete BGZF blocks (thread-safe: no state); blocks compressed elsewhere are appended with write_blocks
    static std::string compress_blocks(const char* p, size_t n, int level) {
        std::string out;
        out.reserve(n / 8 + 64);
        for (size_t a = 0; a < n; a += kBlock) out += compress(p + a, std::min(kBlock, n - a), level);
        return out;
    }
    int level() const { return level_; }
    void write_blocks(const std::string& blocks) {  // whatever text is pending goes out first, as blocks of its own
        flush(true);
        f_.write(blocks.data(), (std::streamsize)blocks.size());
    }
    void finish() {
        flush(true);
        const std::string eof = compress(nullptr, 0, 6);  // EOF marker
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
                out[k] = compress(buf_.data() + k * kBlock, std::min(kBlock, buf_.size() - k * kBlock), level_);
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
    static std::string compress(const char* data, size_t n, int level) {
        std::string out(0x10000 + 64, '\0');
        uint8_t* o = (uint8_t*)&out[0];
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
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
    int level_;
};

}  // namespace
