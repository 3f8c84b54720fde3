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
#include "common.hpp"
#include "options.hpp"
#include "pwm.hpp"
#include "bed.hpp"
#include "bgzf.hpp"
#include "bcf.hpp"
#include "fasta.hpp"
#include "block.hpp"
#include "rows.hpp"

#ifdef TFBS_DRIVER_TEST_SHIM
#include "test_shim.hpp"
#endif

#ifndef TFBS_DRIVER_NO_MAIN
namespace {
#define TF(call)                                                         \
    do {                                                                 \
        int rc_ = (call);                                                \
        if (rc_ != TFBS_OK) die(tfbs_last_error(ctx));                   \
    } while (0)

}  // namespace

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
