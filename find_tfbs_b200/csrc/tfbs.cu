// tfbs.cu -- C ABI (include/tfbs.h) over the sm_100a kernels in kernels.cuh.
//
// One context = one CUDA device + one stream.  A block of merged regions is processed in batches sized to a
// scratch budget: phase 1 groups the haplotypes of every region (K0) and reports per-region group counts,
// phase 2 builds (K1), scans (K2) and reduces (K3) one batch at a time.  There is no CPU fallback: every
// entry point that computes fails with TFBS_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace tfbs;

namespace {

thread_local std::string g_create_error;

// Buffers grow with some slack so that a slightly larger next block does not reallocate; the sanitizer build of tests/cuda_emu sets
// the slack to zero so that every out-of-bounds access is caught.
#ifndef TFBS_ALLOC_SLACK
#define TFBS_ALLOC_SLACK(bytes) ((bytes) / 8 + 256)
#endif

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + TFBS_ALLOC_SLACK(bytes);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {  // pinned
    void* p = nullptr;
    size_t cap = 0;
    ~HostBuf() { if (p) cudaFreeHost(p); }
    cudaError_t reserve(size_t bytes, bool keep) {
        if (bytes <= cap) return cudaSuccess;
        void* np = nullptr;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&np, want);
        if (e != cudaSuccess) return e;
        if (keep && p && cap) memcpy(np, p, cap);
        if (p) cudaFreeHost(p);
        p = np;
        cap = want;
        return cudaSuccess;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct tfbs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string err;

    // options
    int rows_mode = TFBS_ROWS_VARYING;
    int record_matches = 0;
    uint64_t max_matches = 1u << 22;
    int verify_groups = 1;
    int scan_format = 0;
    uint64_t scratch_bytes = 24ull << 30;
    uint32_t table_budget = 96 * 1024;
    int scan_ctas_per_sm = 0;  // < 0: fixed grid size (debugging)
    int delta = 1;             // delta scoring of patched haplotypes
    int64_t refhit_cap_opt = 0; // testing: capacity of the reference-hit buffer (0 = automatic)
    int rows_width = 32;        // 32: counts are returned as u32; 0: narrowest of u8 / u16 / u32 that holds every count of the block
    uint32_t row_bytes = 4;     // element size of the rows held for tfbs_collect
    int audit = 0;              // set by tfbs_audit_block: per-haplotype flags are kept

    // patterns
    bool have_patterns = false;
    std::vector<tfbs_pattern> orig_patterns;          // the caller's list (tfbs_audit_block re-compiles it with lowered thresholds)
    std::vector<std::vector<int32_t>> orig_weights;
    CompiledPatterns cp;
    DevBuf d_table, d_chunks, d_runs, d_trip_pat, d_pat_len, d_pat_pid, d_pid_list;
    DevPatterns dpat{};

    // block inputs (device)
    bool have_block = false;
    uint32_t R = 0, S = 0, H = 0, pitch = 0;
    uint64_t n_ref_bytes = 0, n_allele_bytes = 0, n_var = 0, n_inner = 0, n_carrier_rows = 0;
    DevBuf d_region_start, d_region_end, d_ref_off, d_ref_ascii, d_ref_codes, d_inner_off, d_inner, d_var_off, d_variants,
        d_allele_ascii, d_allele_codes, d_carriers, d_var_class, d_var_inwin, d_ref_prefix;
    // host copies of the small per-region arrays (batch planning, error messages)
    std::vector<int64_t> h_region_start, h_region_end;
    std::vector<uint64_t> h_ref_off;
    std::vector<uint32_t> h_inner_off, h_var_off;
    std::vector<uint64_t> h_ins_extra;  // per region: sum over variants of max(0, alt_len - 1)

    // phase 1
    DevBuf d_sig, d_nd_in, d_leader, d_hap_group, d_ngroups, d_sum_nd, d_ref_used;
    DevBuf d_keys, d_vals;
    std::vector<uint32_t> h_ngroups, h_sum_nd;
    // per region prefix arrays (block-wide)
    std::vector<uint64_t> h_gbase, h_cbase, h_kbase;
    DevBuf d_gbase, d_cbase, d_kbase;

    // phase 2 scratch
    DevBuf d_seq_region, d_seq_leader, d_seq_nd, d_seq_doff, d_dlist, d_segs, d_seq_nseg, d_seq_len, d_ent_units, d_ent_uoff, d_pk,
        d_nm, d_seq_hash, d_seq_flags, d_tile_sums, d_C, d_vmin, d_vmax, d_flag, d_rowidx, d_seq_nitems, d_item_off, d_items, d_refhits,
        d_item_key, d_item_hits, d_item_coff, d_item_cnt, d_score_flag, d_count_size, d_score_idx, d_list, d_refcnt;
    DevBuf d_status;
    DevBuf d_rows_region, d_rows_inner, d_rows_pid, d_rows_vmin, d_rows_vmax, d_rows_left, d_rows_right;
    DevBuf d_m_region, d_m_pattern, d_m_group, d_m_start;
    DevBuf d_hap_flags;

    // results (host, pinned)
    HostBuf h_rows_region, h_rows_inner, h_rows_pid, h_rows_vmin, h_rows_vmax, h_rows_left, h_rows_right;
    HostBuf h_m_region, h_m_pattern, h_m_group, h_m_start, h_hap_group;
    HostBuf h_status, h_totals;
    HostBuf h_hap_flags;
    std::vector<uint32_t> tie_region, tie_pattern, tie_group;  // tfbs_audit_block
    std::vector<int64_t> tie_start;
    uint64_t n_rows = 0, n_matches = 0;
    uint64_t n_matches_found = 0;  // hits the last run produced, also those that did not fit into the match buffer
    bool matches_truncated = false;
    bool ran = false;

    tfbs_stats stats{};
    cudaEvent_t ev[10]{};
};

namespace {

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            ctx->err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;                   \
            return TFBS_ERR_CUDA;                                                                             \
        }                                                                                                     \
    } while (0)

int fail(tfbs_ctx* ctx, int code, const std::string& msg) {
    ctx->err = msg;
    return code;
}

template <class T>
int upload(tfbs_ctx* ctx, DevBuf& buf, const T* src, size_t n) {
    CK(buf.reserve(std::max<size_t>(1, n) * sizeof(T)));
    if (n) CK(cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n * sizeof(T);
    return TFBS_OK;
}

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)std::max<uint64_t>(1, (n + block - 1) / block); }

DevBlock dev_block(const tfbs_ctx* ctx) {
    DevBlock b{};
    b.R = ctx->R;
    b.S = ctx->S;
    b.H = ctx->H;
    b.region_start = ctx->d_region_start.as<i64>();
    b.region_end = ctx->d_region_end.as<i64>();
    b.ref_off = ctx->d_ref_off.as<u64>();
    b.ref_codes = ctx->d_ref_codes.as<u8>();
    b.inner_off = ctx->d_inner_off.as<u32>();
    b.inner = ctx->d_inner.as<tfbs_inner_region>();
    b.var_off = ctx->d_var_off.as<u32>();
    b.variants = ctx->d_variants.as<tfbs_variant>();
    b.allele_codes = ctx->d_allele_codes.as<u8>();
    b.carriers = ctx->d_carriers.as<u32>();
    b.pitch = ctx->pitch;
    b.var_class = ctx->d_var_class.as<u32>();
    b.var_inwin = ctx->d_var_inwin.as<u8>();
    b.ref_prefix = ctx->d_ref_prefix.as<u64>();
    return b;
}

// Shape of the block: everything that is read on the host to size the copies.
int validate_regions(tfbs_ctx* ctx, const tfbs_block* b) {
    if (!b) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "block is NULL");
    if (b->n_regions && (!b->region_start || !b->region_end || !b->ref_off || !b->inner_off || !b->var_off))
        return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "block has NULL region arrays");
    if ((uint64_t)b->n_samples * 2 > 0x7fffffffull) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "too many samples");
    uint32_t H = 2 * b->n_samples;
    if (b->n_carrier_rows && b->carrier_pitch < (H + 31) / 32) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "carrier_pitch is too small");
    for (uint32_t r = 0; r < b->n_regions; ++r) {
        if (b->region_start[r] < 0 || b->region_end[r] < b->region_start[r])
            return fail(ctx, TFBS_ERR_INVALID_ARGUMENT,
                        "region " + std::to_string(r) + " has an invalid extended window (main.rs:407 underflow)");
        if (b->region_end[r] - b->region_start[r] >= (1ll << 26))
            return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "region " + std::to_string(r) + " is longer than 2^26 bases");
        if (b->ref_off[r + 1] < b->ref_off[r] ||
            b->ref_off[r + 1] - b->ref_off[r] > (uint64_t)(b->region_end[r] - b->region_start[r] + 1))
            return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "region " + std::to_string(r) + ": reference window longer than the region");
        if (b->inner_off[r + 1] < b->inner_off[r] || b->var_off[r + 1] < b->var_off[r])
            return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "offset arrays must be non-decreasing");
    }
    return TFBS_OK;
}

// The records: allele ranges and carrier rows (what the kernels index with), and per region the bases insertions can add.
// Runs while the copies of the block are in flight.
int validate_variants(tfbs_ctx* ctx, const tfbs_block* b) {
    ctx->h_ins_extra.assign(b->n_regions, 0);
    for (uint32_t r = 0; r < b->n_regions; ++r)
        for (uint32_t v = b->var_off[r]; v < b->var_off[r + 1]; ++v) {
            const tfbs_variant& x = b->variants[v];
            if (x.ref_len == 0 || x.alt_len == 0) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "variant " + std::to_string(v) + " has an empty allele");
            if ((uint64_t)x.ref_off + x.ref_len > b->allele_bytes || (uint64_t)x.alt_off + x.alt_len > b->allele_bytes)
                return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "variant " + std::to_string(v) + " points outside allele_bases");
            if (x.carrier_row >= b->n_carrier_rows) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "variant " + std::to_string(v) + " has no carrier row");
            if (x.alt_len > 1) ctx->h_ins_extra[r] += x.alt_len - 1;
        }
    return TFBS_OK;
}

int do_upload(tfbs_ctx* ctx, const tfbs_block* b) {
    ctx->have_block = false;
    int rc = validate_regions(ctx, b);
    if (rc) return rc;
    ctx->R = b->n_regions;
    ctx->S = b->n_samples;
    ctx->H = 2 * b->n_samples;
    ctx->pitch = b->carrier_pitch;
    const uint32_t R = ctx->R;
    ctx->n_ref_bytes = R ? b->ref_off[R] : 0;
    ctx->n_inner = R ? b->inner_off[R] : 0;
    ctx->n_var = R ? b->var_off[R] : 0;
    ctx->n_allele_bytes = b->allele_bytes;
    ctx->n_carrier_rows = b->n_carrier_rows;
    ctx->h_region_start.assign(b->region_start, b->region_start + R);
    ctx->h_region_end.assign(b->region_end, b->region_end + R);
    if (R) {
        ctx->h_ref_off.assign(b->ref_off, b->ref_off + R + 1);
        ctx->h_inner_off.assign(b->inner_off, b->inner_off + R + 1);
        ctx->h_var_off.assign(b->var_off, b->var_off + R + 1);
    } else {
        ctx->h_ref_off.assign(1, 0);
        ctx->h_inner_off.assign(1, 0);
        ctx->h_var_off.assign(1, 0);
    }
    if (R && ctx->n_var && !b->variants) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "block has records but variants is NULL");

    if ((rc = upload(ctx, ctx->d_region_start, b->region_start, R))) return rc;
    if ((rc = upload(ctx, ctx->d_region_end, b->region_end, R))) return rc;
    if ((rc = upload(ctx, ctx->d_ref_off, ctx->h_ref_off.data(), R + 1))) return rc;
    if ((rc = upload(ctx, ctx->d_ref_ascii, b->ref_bases, ctx->n_ref_bytes))) return rc;
    if ((rc = upload(ctx, ctx->d_inner_off, ctx->h_inner_off.data(), R + 1))) return rc;
    if ((rc = upload(ctx, ctx->d_inner, b->inner, ctx->n_inner))) return rc;
    if ((rc = upload(ctx, ctx->d_var_off, ctx->h_var_off.data(), R + 1))) return rc;
    if ((rc = upload(ctx, ctx->d_variants, b->variants, ctx->n_var))) return rc;
    if ((rc = upload(ctx, ctx->d_allele_ascii, b->allele_bases, ctx->n_allele_bytes))) return rc;
    if ((rc = upload(ctx, ctx->d_carriers, b->carriers, (size_t)ctx->n_carrier_rows * ctx->pitch))) return rc;
    CK(ctx->d_ref_codes.reserve(std::max<uint64_t>(1, ctx->n_ref_bytes)));
    CK(ctx->d_allele_codes.reserve(std::max<uint64_t>(1, ctx->n_allele_bytes)));
    CK(ctx->d_var_class.reserve(std::max<uint64_t>(1, ctx->n_var) * 4));
    CK(ctx->d_var_inwin.reserve(std::max<uint64_t>(1, ctx->n_var)));
    CK(ctx->d_ref_prefix.reserve((ctx->n_ref_bytes + R + 1) * 8));
    // the per-record checks overlap the copies (asynchronous when the caller's buffers are page-locked); no kernel has been enqueued yet
    if ((rc = validate_variants(ctx, b))) {
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    ctx->have_block = true;
    return TFBS_OK;
}

// ---- the device pipeline on the resident block ---------------------------------------------------
//
// One run = phase 1 over the whole block (input encoding, K0 grouping, per-region prefix arrays), then phase 2 in batches of
// regions sized to the scratch budget: K1 build, K2 work list + scan (+ the finish pass of delta scoring), K3 rows.
struct Pipeline {
    tfbs_ctx* ctx;
    cudaStream_t st;
    const uint32_t R, S, H;
    const uint64_t RH;
    DevStatus* dst = nullptr;
    DevBlock db{};
    DevMatches dm{};
    uint32_t n_pid = 0;
    int smem_bytes = 0;
    bool wide = false;
    uint32_t scan_grid = 0;
    // accumulated over the batches
    float ms_build = 0, ms_scan = 0, ms_count = 0, ms_scan_kernel = 0;
    uint64_t n_items_total = 0;
    uint64_t hits_done = 0;  // hits of the finished batches: what n_hits falls back to when a batch is re-scored in full

    // a batch of regions [r0, r1) and the sizes its scratch arrays are planned for
    struct Batch {
        uint32_t r0 = 0, r1 = 0, nr = 0;
        uint64_t n_seq = 0, n_d = 0, n_units = 0, n_c = 0, n_keys = 0;
        uint64_t items_cap = 0, ic = 1;
        uint32_t capr = 0;  // reference hits kept per region
        DevSeqs sq{};
        DevRefHits drh{};
        DevCounts dc{};
        const u64* d_n_items = nullptr;
        const u64* d_n_list = nullptr;
        uint64_t n_list_host = 0;
        int use_delta = 0;
        DevStatus hs{};  // status word after the batch's rows pass
    };

    explicit Pipeline(tfbs_ctx* c) : ctx(c), st(c->stream), R(c->R), S(c->S), H(c->H), RH((uint64_t)c->R * c->H) {}

    int run() {
        int rc;
        if ((rc = begin())) return rc;
        if (R == 0 || S == 0) {
            CK(cudaStreamSynchronize(st));
            ctx->ran = true;
            return TFBS_OK;
        }
        if ((rc = encode_inputs())) return rc;
        if ((rc = group_haplotypes())) return rc;
        if ((rc = region_prefixes())) return rc;
        if ((rc = setup_scan())) return rc;
        for (uint32_t r0 = 0; r0 < R;) {
            Batch b;
            if ((rc = plan_batch(r0, &b))) return rc;
            if ((rc = reserve_batch(&b))) return rc;
            CK(cudaEventRecord(ctx->ev[2], st));
            if ((rc = build_sequences(b))) return rc;
            CK(cudaEventRecord(ctx->ev[3], st));
            b.use_delta = (ctx->delta && !ctx->record_matches) ? 1 : 0;
            if ((rc = scan_pass(&b, b.use_delta))) return rc;
            CK(cudaEventRecord(ctx->ev[4], st));
            if ((rc = count_and_filter(&b))) return rc;
            if ((rc = fetch_rows(b))) return rc;
            if ((rc = batch_timers())) return rc;
            r0 = b.r1;
        }
        return finish();
    }

private:
    uint32_t& launches() { return ctx->stats.total_launches; }

    // exclusive scan of d_in[0..n) into d_out[0..n], total into d_out[n]
    int device_scan(const uint32_t* d_in, uint64_t n, u64* d_out) {
        uint32_t tiles = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
        if (tiles == 0) {
            CK(cudaMemsetAsync(d_out, 0, sizeof(u64), st));
            return TFBS_OK;
        }
        CK(ctx->d_tile_sums.reserve((size_t)tiles * 8));
        TFBS_LAUNCH(k_prefix_tiles, tiles, SCAN_THREADS, 0, st)(d_in, n, d_out, ctx->d_tile_sums.as<u64>());
        TFBS_LAUNCH(k_prefix_sums, 1, SCAN_THREADS, 0, st)(ctx->d_tile_sums.as<u64>(), tiles, d_out + n);
        TFBS_LAUNCH(k_prefix_add, tiles, SCAN_THREADS, 0, st)(d_out, n, ctx->d_tile_sums.as<u64>());
        launches() += 3;
        CK(cudaGetLastError());
        return TFBS_OK;
    }

    int read_status(DevStatus* hs) {
        CK(cudaMemcpyAsync(ctx->h_status.p, dst, sizeof(DevStatus), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(hs, ctx->h_status.p, sizeof *hs);
        return TFBS_OK;
    }

    int begin() {
        if (!ctx->have_patterns) return fail(ctx, TFBS_ERR_STATE, "tfbs_set_patterns has not been called");
        if (!ctx->have_block) return fail(ctx, TFBS_ERR_STATE, "no block has been uploaded");
        uint64_t h2d_keep = ctx->stats.h2d_bytes;
        memset(&ctx->stats, 0, sizeof ctx->stats);
        ctx->stats.h2d_bytes = h2d_keep;
        ctx->stats.sm_count = (uint32_t)ctx->prop.multiProcessorCount;
        ctx->n_rows = 0;
        ctx->n_matches = 0;
        ctx->matches_truncated = false;
        ctx->ran = false;
        CK(ctx->d_status.reserve(sizeof(DevStatus)));
        CK(ctx->h_status.reserve(sizeof(DevStatus), false));
        CK(ctx->h_totals.reserve(64, false));
        dst = ctx->d_status.as<DevStatus>();
        DevStatus init{};
        init.err_key = ~0ull;
        init.bad_ref_base = ~0ull;
        init.bad_allele_base = ~0ull;
        memcpy(ctx->h_status.p, &init, sizeof init);
        CK(cudaMemcpyAsync(dst, ctx->h_status.p, sizeof init, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ctx->ev[0], st));
        return TFBS_OK;
    }

    // ASCII -> nucleotide codes, Diff classes, prefix hashes of the reference windows
    int encode_inputs() {
        if (ctx->n_ref_bytes) {
            TFBS_LAUNCH(k_encode, std::min<unsigned>(grid_for(ctx->n_ref_bytes, 256), ctx->stats.sm_count * 16), 256, 0, st)(ctx->d_ref_ascii.as<u8>(), ctx->d_ref_codes.as<u8>(),
                                                                                                     ctx->n_ref_bytes, &dst->bad_ref_base);
            ++launches();
        }
        if (ctx->n_allele_bytes) {
            TFBS_LAUNCH(k_encode, std::min<unsigned>(grid_for(ctx->n_allele_bytes, 256), ctx->stats.sm_count * 16), 256, 0, st)(
                ctx->d_allele_ascii.as<u8>(), ctx->d_allele_codes.as<u8>(), ctx->n_allele_bytes, &dst->bad_allele_base);
            ++launches();
        }
        db = dev_block(ctx);
        TFBS_LAUNCH(k_variant_prep, R, 128, 0, st)(db, 0, ctx->d_var_class.as<u32>(), ctx->d_var_inwin.as<u8>());
        TFBS_LAUNCH(k_ref_prefix, R, SCAN_THREADS, 0, st)(db, 0, ctx->d_ref_prefix.as<u64>());
        launches() += 2;
        return TFBS_OK;
    }

    // ---- phase 1: grouping (K0), over super-batches bounded by the hash table; a signature collision retries with another seed ----
    int group_haplotypes() {
        CK(ctx->d_sig.reserve(RH * 8));
        CK(ctx->d_nd_in.reserve(RH * 4));
        CK(ctx->d_leader.reserve(RH * 4));
        CK(ctx->d_hap_group.reserve(RH * 4));
        CK(ctx->d_ngroups.reserve((size_t)R * 4));
        CK(ctx->d_sum_nd.reserve((size_t)R * 4));
        CK(ctx->d_ref_used.reserve((size_t)R * 4));
        if (ctx->audit) CK(ctx->d_hap_flags.reserve(RH));
        const uint64_t max_pairs = 1ull << 25;
        uint32_t regions_per_super = (uint32_t)std::max<uint64_t>(1, max_pairs / std::max<uint32_t>(1, H));
        uint64_t seed = 0x243f6a8885a308d3ull;
        for (int attempt = 0;; ++attempt) {
            for (uint32_t r0 = 0; r0 < R; r0 += regions_per_super) {
                uint32_t nr = std::min(regions_per_super, R - r0);
                uint64_t pairs = (uint64_t)nr * H;
                uint32_t cap = 1024;
                while (cap < 2 * pairs) cap <<= 1;
                CK(ctx->d_keys.reserve((size_t)cap * 8));
                CK(ctx->d_vals.reserve((size_t)cap * 4));
                CK(cudaMemsetAsync(ctx->d_keys.p, 0, (size_t)cap * 8, st));
                CK(cudaMemsetAsync(ctx->d_vals.p, 0xff, (size_t)cap * 4, st));
                TFBS_LAUNCH(k_signatures, grid_for(pairs, 256), 256, 0, st)(db, r0, nr, seed, ctx->d_sig.as<u64>(), ctx->d_nd_in.as<u32>());
                TFBS_LAUNCH(k_group_insert, grid_for(pairs, 256), 256, 0, st)(H, r0, nr, ctx->d_sig.as<u64>(), ctx->d_keys.as<u64>(), ctx->d_vals.as<u32>(), cap - 1);
                TFBS_LAUNCH(k_group_lookup, grid_for(pairs, 256), 256, 0, st)(db, r0, nr, ctx->d_sig.as<u64>(), ctx->d_keys.as<u64>(), ctx->d_vals.as<u32>(), cap - 1,
                                                                     ctx->d_leader.as<u32>(), dst);
                TFBS_LAUNCH(k_group_rank, nr, 256, 0, st)(H, r0, ctx->d_leader.as<u32>(), ctx->d_nd_in.as<u32>(), ctx->d_hap_group.as<u32>(),
                                                 ctx->d_ngroups.as<u32>(), ctx->d_sum_nd.as<u32>());
                launches() += 4;
            }
            CK(cudaGetLastError());
            ctx->h_ngroups.resize(R);
            ctx->h_sum_nd.resize(R);
            CK(cudaMemcpyAsync(ctx->h_ngroups.data(), ctx->d_ngroups.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_sum_nd.data(), ctx->d_sum_nd.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
            DevStatus hs;
            int rc = read_status(&hs);
            if (rc) return rc;
            if (hs.bad_ref_base != ~0ull || hs.bad_allele_base != ~0ull) {
                // util.rs:15 panic!("Unknown nucleotide {}", l)
                return fail(ctx, TFBS_ERR_UNKNOWN_NUCLEOTIDE,
                            std::string("Unknown nucleotide at byte ") +
                                std::to_string(hs.bad_ref_base != ~0ull ? hs.bad_ref_base : hs.bad_allele_base) +
                                (hs.bad_ref_base != ~0ull ? " of ref_bases" : " of allele_bases"));
            }
            if (hs.sig_collision == 0 || !ctx->verify_groups) break;
            if (attempt >= 3) return fail(ctx, TFBS_ERR_INTERNAL, "haplotype signature hash collision persisted over 4 seeds");
            seed = seed * 0x9e3779b97f4a7c15ull + 0x7f4a7c15ull;
            CK(cudaMemsetAsync(&dst->sig_collision, 0, 4, st));
        }
        CK(cudaEventRecord(ctx->ev[1], st));
        return TFBS_OK;
    }

    // per region: first sequence (gbase), first count word (cbase) and first key (kbase), block-wide
    int region_prefixes() {
        n_pid = (uint32_t)ctx->cp.pid_list.size();
        ctx->h_gbase.assign(R + 1, 0);
        ctx->h_cbase.assign(R + 1, 0);
        ctx->h_kbase.assign(R + 1, 0);
        for (uint32_t r = 0; r < R; ++r) {
            uint64_t nk = ctx->h_inner_off[r + 1] - ctx->h_inner_off[r];
            ctx->h_gbase[r + 1] = ctx->h_gbase[r] + ctx->h_ngroups[r];
            ctx->h_cbase[r + 1] = ctx->h_cbase[r] + (uint64_t)ctx->h_ngroups[r] * n_pid * nk;
            ctx->h_kbase[r + 1] = ctx->h_kbase[r] + (uint64_t)n_pid * nk;
        }
        int rc;
        if ((rc = upload(ctx, ctx->d_gbase, ctx->h_gbase.data(), R + 1))) return rc;
        if ((rc = upload(ctx, ctx->d_cbase, ctx->h_cbase.data(), R + 1))) return rc;
        if ((rc = upload(ctx, ctx->d_kbase, ctx->h_kbase.data(), R + 1))) return rc;
        ctx->stats.h2d_bytes -= 3ull * (R + 1) * 8;  // internal traffic, not the caller's inputs
        return TFBS_OK;
    }

    // match buffer, shared-memory size and grid of the scan kernel
    int setup_scan() {
        if (ctx->record_matches) {
            CK(ctx->d_m_region.reserve(ctx->max_matches * 4));
            CK(ctx->d_m_pattern.reserve(ctx->max_matches * 4));
            CK(ctx->d_m_group.reserve(ctx->max_matches * 4));
            CK(ctx->d_m_start.reserve(ctx->max_matches * 8));
        }
        dm.enabled = ctx->record_matches ? 1u : 0u;
        dm.cap = (u32)std::min<uint64_t>(ctx->max_matches, 0xffffffffu);
        dm.region = ctx->d_m_region.as<u32>();
        dm.pattern_index = ctx->d_m_pattern.as<u32>();
        dm.group = ctx->d_m_group.as<u32>();
        dm.start = ctx->d_m_start.as<i64>();
        smem_bytes = (int)(sizeof(CtaShared) + SCAN_WARPS * sizeof(WarpShared) + ((ctx->cp.max_chunk_bytes + 15) & ~15u));
        wide = ctx->cp.fields == 2;
        if (wide) CK(cudaFuncSetAttribute(k_scan<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        else CK(cudaFuncSetAttribute(k_scan<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        scan_grid = (uint32_t)ctx->prop.multiProcessorCount;  // persistent: one CTA per SM shares one copy of the tables
        if (ctx->scan_ctas_per_sm < 0) scan_grid = (uint32_t)std::max(1, -ctx->scan_ctas_per_sm);  // debugging: fixed grid size
        ctx->stats.scan_ctas = scan_grid;
        return TFBS_OK;
    }

    // ---- phase 2: batches under the scratch budget ---------------------------------------------------
    void region_cost(uint32_t r, uint64_t* n_seq, uint64_t* n_d, uint64_t* n_units, uint64_t* n_c, uint64_t* n_keys) const {
        uint64_t g = ctx->h_ngroups[r];
        uint64_t W = (uint64_t)(ctx->h_region_end[r] - ctx->h_region_start[r] + 1) + ctx->h_ins_extra[r];
        uint64_t nk = ctx->h_inner_off[r + 1] - ctx->h_inner_off[r];
        *n_seq = g;
        *n_d = ctx->h_sum_nd[r];
        *n_units = g * ((W + 31) / 32 + 1);
        *n_c = g * n_pid * nk;
        *n_keys = (uint64_t)n_pid * nk;
    }
    static uint64_t bytes_of(uint64_t n_seq, uint64_t n_d, uint64_t n_units, uint64_t n_c, uint64_t n_keys) {
        // sequences, diff lists + segments, packed bases (all of them when delta scoring is off), counts, keys, the sequence-keyed map,
        // and per possible item (<= n_d + n_seq): the item record, key, owner bookkeeping, list entry and its share of the item map
        return n_seq * (4 * 6 + 8 * 3 + 1 + 32) + n_d * (4 + 32) + n_units * 12 + n_c * 4 + n_keys * 24 + n_seq * 2 * 12 * 2 + (n_d + n_seq) * 96;
    }

    int plan_batch(uint32_t r0, Batch* b) {
        b->r0 = r0;
        uint32_t r1 = r0;
        while (r1 < R) {
            uint64_t a, b2, c2, d2, e2;
            region_cost(r1, &a, &b2, &c2, &d2, &e2);
            if (r1 > r0 && (bytes_of(b->n_seq + a, b->n_d + b2, b->n_units + c2, b->n_c + d2, b->n_keys + e2) > ctx->scratch_bytes ||
                            b->n_seq + a > 0x7fffffffull || b->n_seq + a + b->n_d + b2 > (1ull << 30)))
                break;
            b->n_seq += a; b->n_d += b2; b->n_units += c2; b->n_c += d2; b->n_keys += e2;
            ++r1;
        }
        b->r1 = r1;
        b->nr = r1 - r0;
        if (b->n_seq > 0x7fffffffull || b->n_seq + b->n_d > (1ull << 30))
            return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "a single region has more haplotype groups / carried variants than one batch can hold");
        b->items_cap = b->n_d + b->n_seq;
        if (b->items_cap > 0xfffffff0ull) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "batch too large for the scan scheduler");
        b->ic = std::max<uint64_t>(1, b->items_cap);
        // reference hits live in a slab of capr entries per region; a region with more hits sends the batch to the full scan
        b->capr = ctx->refhit_cap_opt ? (uint32_t)std::max<int64_t>(1, ctx->refhit_cap_opt / std::max<uint32_t>(1, b->nr))
                                      : (uint32_t)std::min<uint64_t>(32768, std::max<uint64_t>(512, (1ull << 30) / std::max<uint32_t>(1, b->nr) / sizeof(RefHit)));  // <= 1 GB of slabs
        return TFBS_OK;
    }

    int reserve_batch(Batch* b) {
        const uint64_t n_seq = b->n_seq, n_d = b->n_d, n_c = b->n_c, n_keys = b->n_keys, ic = b->ic;
        CK(ctx->d_seq_region.reserve(n_seq * 4));
        CK(ctx->d_seq_leader.reserve(n_seq * 4));
        CK(ctx->d_seq_nd.reserve(n_seq * 4));
        CK(ctx->d_seq_doff.reserve((n_seq + 1) * 8));
        CK(ctx->d_dlist.reserve(std::max<uint64_t>(1, n_d) * 4));
        CK(ctx->d_segs.reserve((2 * n_d + 2 * n_seq) * sizeof(Seg)));
        CK(ctx->d_seq_nseg.reserve(n_seq * 4));
        CK(ctx->d_seq_len.reserve(n_seq * 4));
        CK(ctx->d_seq_hash.reserve(n_seq * 8));
        CK(ctx->d_seq_flags.reserve(n_seq));
        CK(ctx->d_C.reserve(std::max<uint64_t>(1, n_c) * 4));
        CK(ctx->d_vmin.reserve(std::max<uint64_t>(1, n_keys) * 4));
        CK(ctx->d_vmax.reserve(std::max<uint64_t>(1, n_keys) * 4));
        CK(ctx->d_flag.reserve(std::max<uint64_t>(1, n_keys) * 4));
        CK(ctx->d_rowidx.reserve((n_keys + 1) * 8));
        CK(ctx->d_refcnt.reserve((size_t)b->nr * 4));
        CK(ctx->d_seq_nitems.reserve(n_seq * 4));
        CK(ctx->d_item_off.reserve((n_seq + 1) * 8));
        CK(ctx->d_items.reserve(ic * sizeof(ScanItem)));
        CK(ctx->d_refhits.reserve((size_t)b->capr * b->nr * sizeof(RefHit)));
        CK(ctx->d_item_key.reserve(ic * 8));
        CK(ctx->d_item_hits.reserve(ic * 4));
        CK(ctx->d_item_coff.reserve((ic + 1) * 8));
        CK(ctx->d_score_flag.reserve(ic * 4));
        CK(ctx->d_count_size.reserve(ic * 4));
        CK(ctx->d_score_idx.reserve((ic + 1) * 8));
        CK(ctx->d_list.reserve(ic * 4));
        CK(ctx->d_ent_units.reserve(ic * 4));
        CK(ctx->d_ent_uoff.reserve((ic + 1) * 8));

        DevSeqs& sq = b->sq;
        sq.n_seq = (u32)n_seq;
        sq.gbase = ctx->d_gbase.as<u64>();
        sq.gbase0 = ctx->h_gbase[b->r0];
        sq.seq_region = ctx->d_seq_region.as<u32>();
        sq.seq_leader = ctx->d_seq_leader.as<u32>();
        sq.seq_nd = ctx->d_seq_nd.as<u32>();
        sq.seq_doff = ctx->d_seq_doff.as<u64>();
        sq.dlist = ctx->d_dlist.as<u32>();
        sq.segs = ctx->d_segs.as<Seg>();
        sq.seq_nseg = ctx->d_seq_nseg.as<u32>();
        sq.seq_len = ctx->d_seq_len.as<u32>();
        sq.ent_units = ctx->d_ent_units.as<u32>();
        sq.ent_uoff = ctx->d_ent_uoff.as<u64>();
        sq.pk = nullptr;  // sized once the scored entries are known
        sq.nm = nullptr;
        sq.seq_hash = ctx->d_seq_hash.as<u64>();
        sq.seq_flags = ctx->d_seq_flags.as<u8>();
        sq.seq_nitems = ctx->d_seq_nitems.as<u32>();
        sq.item_off = ctx->d_item_off.as<u64>();
        sq.items = ctx->d_items.as<ScanItem>();
        sq.n_items_cap = (u32)std::min<uint64_t>(b->items_cap, 0xffffffffu);
        sq.item_key = ctx->d_item_key.as<u64>();
        sq.item_hits = ctx->d_item_hits.as<u32>();
        sq.item_coff = ctx->d_item_coff.as<u64>();
        sq.item_cnt = nullptr;  // sized once the owners are known
        b->drh = DevRefHits{ctx->d_refhits.as<RefHit>(), ctx->d_refcnt.as<u32>(), b->capr, b->r0};
        b->dc.C = ctx->d_C.as<u32>();
        b->dc.cbase = ctx->d_cbase.as<u64>();
        b->dc.cbase0 = ctx->h_cbase[b->r0];
        b->d_n_items = sq.item_off + n_seq;
        b->d_n_list = ctx->d_score_idx.as<u64>() + b->items_cap;
        return TFBS_OK;
    }

    // K1: segments + hash of every distinct haplotype, then the sequence-keyed map of load_haplotypes
    int build_sequences(Batch& b) {
        int rc;
        const uint64_t n_seq = b.n_seq;
        TFBS_LAUNCH(k_seq_init, b.nr, 128, 0, st)(H, b.r0, ctx->d_hap_group.as<u32>(), ctx->d_leader.as<u32>(), ctx->d_nd_in.as<u32>(), b.sq);
        ++launches();
        if ((rc = device_scan(b.sq.seq_nd, n_seq, b.sq.seq_doff))) return rc;
        TFBS_LAUNCH(k_walk, grid_for(n_seq, 128), 128, 0, st)(db, b.sq, dst);
        ++launches();
        uint32_t cap = 1024;
        while (cap < 2 * n_seq) cap <<= 1;
        CK(ctx->d_keys.reserve((size_t)cap * 8));
        CK(ctx->d_vals.reserve((size_t)cap * 4));
        CK(cudaMemsetAsync(ctx->d_keys.p, 0, (size_t)cap * 8, st));
        CK(cudaMemsetAsync(ctx->d_vals.p, 0xff, (size_t)cap * 4, st));
        CK(cudaMemsetAsync(ctx->d_ref_used.as<u32>() + b.r0, 0, (size_t)b.nr * 4, st));
        TFBS_LAUNCH(k_seq_insert, grid_for(n_seq, 256), 256, 0, st)(b.sq, ctx->d_keys.as<u64>(), ctx->d_vals.as<u32>(), cap - 1);
        TFBS_LAUNCH(k_seq_resolve, grid_for(n_seq, 128), 128, 0, st)(db, b.sq, ctx->d_keys.as<u64>(), ctx->d_vals.as<u32>(), cap - 1, dst);
        TFBS_LAUNCH(k_redirect, grid_for((uint64_t)b.nr * H, 256), 256, 0, st)(H, b.r0, b.nr, b.sq, ctx->d_hap_group.as<u32>(), ctx->d_ref_used.as<u32>(),
                                                                        ctx->audit ? ctx->d_hap_flags.as<u8>() : nullptr);
        launches() += 3;
        return TFBS_OK;
    }

    // K2: work list, packing of the scored bases, the scan (one launch per pattern chunk), the finish pass of delta scoring
    int scan_pass(Batch* bp, int delta) {
        Batch& b = *bp;
        DevSeqs& sq = b.sq;
        int rc;
        const uint64_t n_seq = b.n_seq, items_cap = b.items_cap, ic = b.ic;
        if (b.n_c) CK(cudaMemsetAsync(ctx->d_C.p, 0, b.n_c * 4, st));
        CK(cudaMemsetAsync(ctx->d_refcnt.p, 0, (size_t)b.nr * 4, st));
        uint32_t tcap = 1024;
        if (delta) {
            while (tcap < 2 * items_cap) tcap <<= 1;
            CK(ctx->d_keys.reserve((size_t)tcap * 8));
            CK(ctx->d_vals.reserve((size_t)tcap * 4));
            CK(cudaMemsetAsync(ctx->d_keys.p, 0, (size_t)tcap * 8, st));
            CK(cudaMemsetAsync(ctx->d_vals.p, 0xff, (size_t)tcap * 4, st));
        }
        CK(cudaMemsetAsync(ctx->d_score_flag.p, 0, ic * 4, st));
        CK(cudaMemsetAsync(ctx->d_count_size.p, 0, ic * 4, st));
        CK(cudaMemsetAsync(ctx->d_ent_units.p, 0, ic * 4, st));
        TFBS_LAUNCH(k_items<false>, grid_for(n_seq, 128), 128, 0, st)(db, sq, ctx->d_ref_used.as<u32>(), ctx->cp.max_len, delta, ctx->d_keys.as<u64>(),
                                                            ctx->d_vals.as<u32>(), tcap - 1);
        ++launches();
        if ((rc = device_scan(sq.seq_nitems, n_seq, sq.item_off))) return rc;
        TFBS_LAUNCH(k_items<true>, grid_for(n_seq, 128), 128, 0, st)(db, sq, ctx->d_ref_used.as<u32>(), ctx->cp.max_len, delta, ctx->d_keys.as<u64>(),
                                                           ctx->d_vals.as<u32>(), tcap - 1);
        TFBS_LAUNCH(k_item_resolve, grid_for(items_cap, 128), 128, 0, st)(db, sq, ctx->dpat, b.d_n_items, delta, ctx->cp.max_len, ctx->d_keys.as<u64>(),
                                                                ctx->d_vals.as<u32>(), tcap - 1, ctx->d_score_flag.as<u32>(),
                                                                ctx->d_count_size.as<u32>());
        launches() += 2;
        if ((rc = device_scan(ctx->d_score_flag.as<u32>(), items_cap, ctx->d_score_idx.as<u64>()))) return rc;
        if ((rc = device_scan(ctx->d_count_size.as<u32>(), items_cap, sq.item_coff))) return rc;
        TFBS_LAUNCH(k_item_lists, grid_for(items_cap, 256), 256, 0, st)(sq, b.d_n_items, ctx->d_score_flag.as<u32>(), ctx->d_score_idx.as<u64>(),
                                                              ctx->d_list.as<u32>());
        ++launches();
        if ((rc = device_scan(sq.ent_units, items_cap, sq.ent_uoff))) return rc;
        // the sizes of the owners' count vectors and of the packed bases decide two allocations: one round trip to the host
        CK(cudaMemcpyAsync((char*)ctx->h_totals.p + 16, sq.item_coff + items_cap, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync((char*)ctx->h_totals.p + 24, b.d_n_list, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync((char*)ctx->h_totals.p + 32, sq.ent_uoff + items_cap, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint64_t cnt_words = ctx->h_totals.as<uint64_t>()[2];
        b.n_list_host = ctx->h_totals.as<uint64_t>()[3];
        const uint64_t ent_units_total = ctx->h_totals.as<uint64_t>()[4];
        CK(ctx->d_item_cnt.reserve(std::max<uint64_t>(1, cnt_words) * 4));
        if (cnt_words) CK(cudaMemsetAsync(ctx->d_item_cnt.p, 0, cnt_words * 4, st));
        sq.item_cnt = ctx->d_item_cnt.as<u32>();
        CK(ctx->d_pk.reserve(std::max<uint64_t>(1, ent_units_total) * 8));
        CK(ctx->d_nm.reserve(std::max<uint64_t>(1, ent_units_total) * 4));
        sq.pk = ctx->d_pk.as<u64>();
        sq.nm = ctx->d_nm.as<u32>();
        if (b.n_list_host) {
            TFBS_LAUNCH(k_emit_list, grid_for(b.n_list_host * EMIT_LANES, 256), 256, 0, st)(db, sq, ctx->d_list.as<u32>(), b.d_n_list);
            TFBS_LAUNCH(k_item_stats, grid_for(b.n_list_host, 256), 256, 0, st)(sq, ctx->dpat, ctx->d_list.as<u32>(), b.d_n_list, dst);
            launches() += 2;
        }
        const u32 per_grab = delta ? (u32)SCAN_PER_GRAB : 1u;  // short items: several list entries per trip to the work counter
        CK(cudaEventRecord(ctx->ev[8], st));
        for (uint32_t c = 0; c < ctx->cp.chunks.size() && b.n_list_host; ++c) {
            CK(cudaMemsetAsync(&dst->work_counter, 0, 4, st));
            if (wide) TFBS_LAUNCH(k_scan<2>, scan_grid, SCAN_CTA, smem_bytes, st)(db, sq, ctx->dpat, b.dc, dm, b.drh, ctx->d_list.as<u32>(), b.d_n_list, per_grab, dst, c, delta);
            else TFBS_LAUNCH(k_scan<3>, scan_grid, SCAN_CTA, smem_bytes, st)(db, sq, ctx->dpat, b.dc, dm, b.drh, ctx->d_list.as<u32>(), b.d_n_list, per_grab, dst, c, delta);
            ++launches();
            ++ctx->stats.scan_launches;
            ctx->stats.scan_input_bytes += ent_units_total * 12 + (uint64_t)ctx->cp.chunks[c].tbl_words * 8 * scan_grid;
        }
        CK(cudaEventRecord(ctx->ev[9], st));
        if (delta) {
            TFBS_LAUNCH(k_group_finish, grid_for(n_seq * 8, 256), 256, 0, st)(db, sq, ctx->dpat, b.dc, b.drh, ctx->d_ref_used.as<u32>(), dst);
            ++launches();
        }
        CK(cudaGetLastError());
        return TFBS_OK;
    }

    // min / max per key and the row index of every emitted key; the row count lands in h_totals[0]
    int rows_pass(Batch& b, int delta) {
        if (!b.n_keys) return TFBS_OK;
        int rc;
        TFBS_LAUNCH(k_rows_minmax, b.nr, 128, 0, st)(db, b.r0, ctx->d_hap_group.as<u32>(), b.dc, ctx->d_gbase.as<u64>(), n_pid, ctx->d_kbase.as<u64>(),
                                          ctx->h_kbase[b.r0], ctx->rows_mode, delta, ctx->d_vmin.as<u32>(), ctx->d_vmax.as<u32>(), ctx->d_flag.as<u32>(),
                                          &dst->max_count);
        ++launches();
        if ((rc = device_scan(ctx->d_flag.as<u32>(), b.n_keys, ctx->d_rowidx.as<u64>()))) return rc;
        CK(cudaMemcpyAsync(ctx->h_totals.p, ctx->d_rowidx.as<u64>() + b.n_keys, 8, cudaMemcpyDeviceToHost, st));
        return TFBS_OK;
    }

    // K3 up to the row count; reports the reference's panics; re-scores the batch in full when the reference-hit slabs overflowed
    int count_and_filter(Batch* bp) {
        Batch& b = *bp;
        int rc;
        TFBS_LAUNCH(k_nominal, grid_for((uint64_t)b.nr * H, 256), 256, 0, st)(db, b.r0, b.nr, ctx->d_hap_group.as<u32>(), b.sq, ctx->dpat, dst);
        TFBS_LAUNCH(k_seq_stats, grid_for(b.n_seq, 256), 256, 0, st)(b.sq, ctx->dpat, ctx->d_ref_used.as<u32>(), dst);
        launches() += 2;
        if ((rc = rows_pass(b, b.use_delta))) return rc;
        if ((rc = read_status(&b.hs))) return rc;
        CK(cudaGetLastError());
        const DevStatus& hs = b.hs;
        if (hs.err_key != ~0ull) {
            uint32_t q = (uint32_t)(hs.err_key >> 32);
            int64_t rel = (int64_t)((hs.err_key >> 4) & 0xfffffff) - (1 << 27);
            uint32_t code = (uint32_t)(hs.err_key & 15);
            // region of sequence q: last r with gbase[r] - gbase[r0] <= q
            uint32_t r = (uint32_t)(std::upper_bound(ctx->h_gbase.begin() + b.r0, ctx->h_gbase.begin() + b.r1, ctx->h_gbase[b.r0] + q) - ctx->h_gbase.begin() - 1);
            int64_t pos = ctx->h_region_start[r] + rel;
            if (code == DEV_REF_MISMATCH)
                return fail(ctx, TFBS_ERR_REF_MISMATCH,
                            "First reference nucleotide of variant doesn't match reference genome: ref_position=" + std::to_string(pos) +
                                " region=" + std::to_string(r));
            return fail(ctx, TFBS_ERR_MISSING_CASE, "Missing case in haplotype patcher (ref_position=" + std::to_string(pos) + " region=" + std::to_string(r) + ")");
        }
        if (hs.seq_collision) return fail(ctx, TFBS_ERR_INTERNAL, "sequence hash collision between distinct haplotypes");
        if (b.use_delta && hs.refhit_overflow) {
            // more reference hits than the slabs hold (very permissive thresholds): score this batch in full instead
            DevStatus fix = hs;
            fix.refhit_overflow = 0;
            fix.n_refhits = 0;
            fix.n_hits = hits_done;
            memcpy(ctx->h_status.p, &fix, sizeof fix);
            CK(cudaMemcpyAsync(dst, ctx->h_status.p, sizeof fix, cudaMemcpyHostToDevice, st));
            b.use_delta = 0;
            if ((rc = scan_pass(bp, 0))) return rc;
            if ((rc = rows_pass(b, 0))) return rc;
            if ((rc = read_status(&b.hs))) return rc;
            CK(cudaGetLastError());
        }
        hits_done = b.hs.n_hits;
        n_items_total += b.n_list_host;
        return TFBS_OK;
    }

    // compaction of the emitted rows and their copy into the pinned result buffers (appended to the rows of earlier batches)
    int fetch_rows(Batch& b) {
        const uint64_t batch_rows = b.n_keys ? *ctx->h_totals.as<uint64_t>() : 0;
        if (batch_rows) {
            const uint64_t n_keys = b.n_keys;
            uint64_t tot = ctx->n_rows + batch_rows;
            // element width of left / right: u32 like the reference's Vec<u32>, or (option rows_width = 0) the narrowest type that
            // holds every count of the block: the rows are the dominant PCIe traffic of large cohorts
            uint32_t eb = 4;
            if (ctx->rows_width == 0) eb = b.hs.max_count < 256 ? 1 : (b.hs.max_count < 65536 ? 2 : 4);
            if (ctx->n_rows == 0) ctx->row_bytes = eb;
            if (eb > ctx->row_bytes) {  // an earlier batch of this block was stored narrower: widen it in place (rare)
                const uint64_t n = ctx->n_rows * S;
                CK(ctx->h_rows_left.reserve(tot * S * eb, true));
                CK(ctx->h_rows_right.reserve(tot * S * eb, true));
                for (HostBuf* hb : {&ctx->h_rows_left, &ctx->h_rows_right})
                    for (uint64_t i = n; i-- > 0;) {
                        uint32_t v = ctx->row_bytes == 1 ? hb->as<uint8_t>()[i] : hb->as<uint16_t>()[i];
                        if (eb == 2) hb->as<uint16_t>()[i] = (uint16_t)v; else hb->as<uint32_t>()[i] = v;
                    }
                ctx->row_bytes = eb;
            }
            eb = ctx->row_bytes;
            CK(ctx->d_rows_region.reserve(batch_rows * 4));
            CK(ctx->d_rows_inner.reserve(batch_rows * 4));
            CK(ctx->d_rows_pid.reserve(batch_rows * 2));
            CK(ctx->d_rows_vmin.reserve(batch_rows * 4));
            CK(ctx->d_rows_vmax.reserve(batch_rows * 4));
            CK(ctx->d_rows_left.reserve(batch_rows * S * eb));
            CK(ctx->d_rows_right.reserve(batch_rows * S * eb));
            CK(ctx->h_rows_region.reserve(tot * 4, true));
            CK(ctx->h_rows_inner.reserve(tot * 4, true));
            CK(ctx->h_rows_pid.reserve(tot * 2, true));
            CK(ctx->h_rows_vmin.reserve(tot * 4, true));
            CK(ctx->h_rows_vmax.reserve(tot * 4, true));
            CK(ctx->h_rows_left.reserve(tot * S * eb, true));
            CK(ctx->h_rows_right.reserve(tot * S * eb, true));
            DevRows dr{ctx->d_rows_region.as<u32>(), ctx->d_rows_inner.as<u32>(), ctx->d_rows_pid.as<u16>(), ctx->d_rows_vmin.as<u32>(),
                       ctx->d_rows_vmax.as<u32>(), ctx->d_rows_left.p, ctx->d_rows_right.p};
#define TFBS_ROWS_WRITE(T)                                                                                                              \
    TFBS_LAUNCH(k_rows_write<T>, grid_for(n_keys * 32, 256), 256, 0, st)(db, b.r0, b.nr, ctx->d_hap_group.as<u32>(), b.dc, n_pid, ctx->d_pid_list.as<u16>(), \
                                                                ctx->d_kbase.as<u64>(), ctx->h_kbase[b.r0], n_keys, ctx->d_vmin.as<u32>(),   \
                                                                ctx->d_vmax.as<u32>(), ctx->d_flag.as<u32>(), ctx->d_rowidx.as<u64>(), dr, 0, b.use_delta)
            if (eb == 1) TFBS_ROWS_WRITE(u8);
            else if (eb == 2) TFBS_ROWS_WRITE(u16);
            else TFBS_ROWS_WRITE(u32);
#undef TFBS_ROWS_WRITE
            ++launches();
            uint64_t o = ctx->n_rows;
            CK(cudaMemcpyAsync(ctx->h_rows_region.as<u32>() + o, dr.region, batch_rows * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_rows_inner.as<u32>() + o, dr.inner, batch_rows * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_rows_pid.as<u16>() + o, dr.pattern_id, batch_rows * 2, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_rows_vmin.as<u32>() + o, dr.vmin, batch_rows * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_rows_vmax.as<u32>() + o, dr.vmax, batch_rows * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_rows_left.as<uint8_t>() + o * S * eb, dr.left, batch_rows * S * eb, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->h_rows_right.as<uint8_t>() + o * S * eb, dr.right, batch_rows * S * eb, cudaMemcpyDeviceToHost, st));
            ctx->stats.d2h_bytes += batch_rows * (4 * 4 + 2 + 2ull * eb * S);
            ctx->n_rows = tot;
        }
        CK(cudaEventRecord(ctx->ev[5], st));
        CK(cudaStreamSynchronize(st));
        return TFBS_OK;
    }

    int batch_timers() {
        float t;
        CK(cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3]));
        ms_build += t;
        CK(cudaEventElapsedTime(&t, ctx->ev[3], ctx->ev[4]));
        ms_scan += t;
        CK(cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]));
        ms_count += t;
        CK(cudaEventElapsedTime(&t, ctx->ev[8], ctx->ev[9]));
        ms_scan_kernel += t;
        return TFBS_OK;
    }

    // final status word, the match list and the audit flags, timings and counters of the run
    int finish() {
        CK(cudaEventRecord(ctx->ev[6], st));
        if (ctx->record_matches) {
            CK(ctx->h_hap_group.reserve(RH * 4, false));
            CK(cudaMemcpyAsync(ctx->h_hap_group.p, ctx->d_hap_group.p, RH * 4, cudaMemcpyDeviceToHost, st));
        }
        if (ctx->audit) {
            CK(ctx->h_hap_flags.reserve(RH, false));
            CK(cudaMemcpyAsync(ctx->h_hap_flags.p, ctx->d_hap_flags.p, RH, cudaMemcpyDeviceToHost, st));
        }
        DevStatus hs;
        int rc = read_status(&hs);
        if (rc) return rc;
        if (ctx->record_matches) {
            uint64_t n = std::min<uint64_t>(hs.n_matches, dm.cap);
            ctx->matches_truncated = hs.n_matches > dm.cap;
            ctx->n_matches_found = hs.n_matches;
            CK(ctx->h_m_region.reserve(std::max<uint64_t>(1, n) * 4, false));
            CK(ctx->h_m_pattern.reserve(std::max<uint64_t>(1, n) * 4, false));
            CK(ctx->h_m_group.reserve(std::max<uint64_t>(1, n) * 4, false));
            CK(ctx->h_m_start.reserve(std::max<uint64_t>(1, n) * 8, false));
            if (n) {
                CK(cudaMemcpyAsync(ctx->h_m_region.p, dm.region, n * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(ctx->h_m_pattern.p, dm.pattern_index, n * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(ctx->h_m_group.p, dm.group, n * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(ctx->h_m_start.p, dm.start, n * 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
            }
            ctx->n_matches = n;
        }
        float t;
        CK(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
        ctx->stats.ms_group = t;
        CK(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[6]));
        ctx->stats.ms_total = t;
        ctx->stats.ms_build = ms_build;
        ctx->stats.ms_scan = ms_scan;
        ctx->stats.ms_count = ms_count;
        ctx->stats.n_regions = R;
        ctx->stats.n_groups = hs.n_scanned;
        ctx->stats.executed_cells = hs.executed_cells;
        ctx->stats.nominal_cells = hs.nominal_cells;
        ctx->stats.n_hits = hs.n_hits;
        ctx->stats.n_keys = ctx->h_kbase[R];
        ctx->stats.n_rows = ctx->n_rows;
        ctx->stats.evaluated_cells = hs.evaluated_cells;
        ctx->stats.n_scan_items = n_items_total;
        ctx->stats.ms_scan_kernel = ms_scan_kernel;
        ctx->stats.n_dropped = hs.n_dropped;
        ctx->stats.n_truncated = hs.n_truncated;
        ctx->ran = true;
        return TFBS_OK;
    }
};

int run_pipeline(tfbs_ctx* ctx) { return Pipeline(ctx).run(); }


}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int tfbs_abi_version(void) { return TFBS_ABI_VERSION; }

int tfbs_create(int device, tfbs_ctx** out) {
    if (!out) return TFBS_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        return TFBS_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        g_create_error = "device index out of range";
        return TFBS_ERR_INVALID_ARGUMENT;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice failed: ") + cudaGetErrorString(e);
        return TFBS_ERR_CUDA;
    }
    tfbs_ctx* ctx = new tfbs_ctx();
    ctx->device = device;
    if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
        delete ctx;
        return TFBS_ERR_CUDA;
    }
    if (ctx->prop.major < 10) {
        g_create_error = std::string("device ") + ctx->prop.name + " is not sm_100-class; this library carries sm_100a code only";
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return TFBS_ERR_CUDA;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    *out = ctx;
    return TFBS_OK;
}

void tfbs_destroy(tfbs_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* tfbs_last_error(const tfbs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tfbs_set_option(tfbs_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return TFBS_ERR_INVALID_ARGUMENT;
    std::string k(key);
    if (k == "rows_mode") {
        if (value != TFBS_ROWS_VARYING && value != TFBS_ROWS_ALL_KEYS) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "rows_mode must be 0 or 1");
        ctx->rows_mode = (int)value;
    } else if (k == "record_matches") ctx->record_matches = value != 0;
    else if (k == "max_matches") ctx->max_matches = (uint64_t)std::max<int64_t>(1, value);
    else if (k == "verify_groups") ctx->verify_groups = value != 0;
    else if (k == "scan_format") { ctx->scan_format = (int)value; ctx->have_patterns = false; }
    else if (k == "scratch_mb") ctx->scratch_bytes = (uint64_t)std::max<int64_t>(64, value) << 20;
    else if (k == "table_budget_kb") { ctx->table_budget = (uint32_t)std::max<int64_t>(8, value) * 1024; ctx->have_patterns = false; }
    else if (k == "scan_ctas_per_sm") ctx->scan_ctas_per_sm = (int)value;
    else if (k == "delta") ctx->delta = value != 0;
    else if (k == "refhit_cap") ctx->refhit_cap_opt = std::max<int64_t>(0, value);
    else if (k == "rows_width") {
        if (value != 0 && value != 32) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "rows_width must be 32 or 0 (automatic)");
        ctx->rows_width = (int)value;
    }
    else return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "unknown option " + k);
    return TFBS_OK;
}

// Compile the pattern list into scan tables and upload them.
static int install_patterns(tfbs_ctx* ctx, const tfbs_pattern* patterns, uint32_t n_patterns) {
    CK(cudaSetDevice(ctx->device));
    ctx->have_patterns = false;
    if (n_patterns == 0) return fail(ctx, TFBS_ERR_INVALID_ARGUMENT, "assertion failed: pwm_list.len() > 0");  // main.rs:238
    size_t max_smem = ctx->prop.sharedMemPerBlockOptin;
    size_t fixed = sizeof(CtaShared) + SCAN_WARPS * sizeof(WarpShared) + 1024;
    uint32_t budget = (uint32_t)std::min<size_t>(ctx->table_budget, max_smem > fixed ? max_smem - fixed : 0);
    std::string err;
    CompiledPatterns cp;
    int rc = compile_patterns(patterns, n_patterns, budget, ctx->scan_format == 1, &cp, &err);
    if (rc != TFBS_OK) return fail(ctx, rc, err);
    ctx->cp = std::move(cp);
    const CompiledPatterns& c = ctx->cp;
    uint64_t keep = ctx->stats.h2d_bytes;
    std::vector<uint64_t> table = c.table;
    table.resize(table.size() + 2, 0);  // 128-bit loads may read one word past the end
    if ((rc = upload(ctx, ctx->d_table, table.data(), table.size()))) return rc;
    if ((rc = upload(ctx, ctx->d_chunks, c.chunks.data(), c.chunks.size()))) return rc;
    if ((rc = upload(ctx, ctx->d_runs, c.runs.data(), c.runs.size()))) return rc;
    if ((rc = upload(ctx, ctx->d_trip_pat, c.trip_pat.data(), c.trip_pat.size()))) return rc;
    if ((rc = upload(ctx, ctx->d_pat_len, c.pat_len.data(), c.pat_len.size()))) return rc;
    if ((rc = upload(ctx, ctx->d_pat_pid, c.pat_pid_index.data(), c.pat_pid_index.size()))) return rc;
    if ((rc = upload(ctx, ctx->d_pid_list, c.pid_list.data(), c.pid_list.size()))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes = keep;
    DevPatterns& d = ctx->dpat;
    d.table = ctx->d_table.as<u64>();
    d.chunks = ctx->d_chunks.as<ChunkDesc>();
    d.runs = ctx->d_runs.as<RunDesc>();
    d.trip_pat = ctx->d_trip_pat.as<int>();
    d.pat_len = ctx->d_pat_len.as<u32>();
    d.pat_pid_index = ctx->d_pat_pid.as<u32>();
    d.n_chunks = (u32)c.chunks.size();
    d.n_pid = (u32)c.pid_list.size();
    d.n_patterns = (u32)c.patterns.size();
    d.max_len = c.max_len;
    d.sum_len = c.sum_len;
    d.sum_len_sq = c.sum_len_sq;
    ctx->have_patterns = true;
    return TFBS_OK;
}

int tfbs_set_patterns(tfbs_ctx* ctx, const tfbs_pattern* patterns, uint32_t n_patterns) {
    if (!ctx || (!patterns && n_patterns)) return TFBS_ERR_INVALID_ARGUMENT;
    int rc = install_patterns(ctx, patterns, n_patterns);
    if (rc != TFBS_OK) return rc;
    // keep the caller's list: tfbs_audit_block compiles it a second time with every threshold lowered by one
    ctx->orig_patterns.assign(patterns, patterns + n_patterns);
    ctx->orig_weights.assign(n_patterns, std::vector<int32_t>());
    for (uint32_t i = 0; i < n_patterns; ++i)
        if (patterns[i].kind == TFBS_PATTERN_PWM && patterns[i].weights && patterns[i].len) {
            ctx->orig_weights[i].assign(patterns[i].weights, patterns[i].weights + 4 * (size_t)patterns[i].len);
            ctx->orig_patterns[i].weights = ctx->orig_weights[i].data();
        } else {
            ctx->orig_patterns[i].weights = nullptr;
        }
    return TFBS_OK;
}

int tfbs_upload_block(tfbs_ctx* ctx, const tfbs_block* block) {
    if (!ctx) return TFBS_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    ctx->stats.h2d_bytes = 0;
    int rc = do_upload(ctx, block);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return TFBS_OK;
}

int tfbs_run_resident(tfbs_ctx* ctx) {
    if (!ctx) return TFBS_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    return run_pipeline(ctx);
}

int tfbs_submit_block(tfbs_ctx* ctx, const tfbs_block* block) {
    if (!ctx) return TFBS_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->have_patterns) return fail(ctx, TFBS_ERR_STATE, "tfbs_set_patterns has not been called");
    ctx->stats.h2d_bytes = 0;
    int rc = do_upload(ctx, block);
    if (rc) return rc;
    return run_pipeline(ctx);
}

int tfbs_collect(tfbs_ctx* ctx, tfbs_rows* out) {
    if (!ctx || !out) return TFBS_ERR_INVALID_ARGUMENT;
    if (!ctx->ran) return fail(ctx, TFBS_ERR_STATE, "tfbs_collect called before a successful tfbs_submit_block / tfbs_run_resident");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    out->n_rows = ctx->n_rows;
    out->n_samples = ctx->S;
    out->count_bytes = ctx->n_rows ? ctx->row_bytes : 4;
    out->region = ctx->h_rows_region.as<uint32_t>();
    out->inner = ctx->h_rows_inner.as<uint32_t>();
    out->pattern_id = ctx->h_rows_pid.as<uint16_t>();
    out->vmin = ctx->h_rows_vmin.as<uint32_t>();
    out->vmax = ctx->h_rows_vmax.as<uint32_t>();
    out->left = ctx->h_rows_left.as<uint32_t>();
    out->right = ctx->h_rows_right.as<uint32_t>();
    return TFBS_OK;
}

int tfbs_get_matches(tfbs_ctx* ctx, tfbs_matches* out) {
    if (!ctx || !out) return TFBS_ERR_INVALID_ARGUMENT;
    if (!ctx->ran || !ctx->record_matches) return fail(ctx, TFBS_ERR_STATE, "matches were not recorded (set option record_matches before the run)");
    out->n_matches = ctx->n_matches;
    out->region = ctx->h_m_region.as<uint32_t>();
    out->pattern_index = ctx->h_m_pattern.as<uint32_t>();
    out->group = ctx->h_m_group.as<uint32_t>();
    out->start = ctx->h_m_start.as<int64_t>();
    out->hap_group = ctx->h_hap_group.as<uint32_t>();
    out->n_samples = ctx->S;
    out->truncated = ctx->matches_truncated ? 1 : 0;
    return TFBS_OK;
}

// Ties = windows reported with min_score - 1 but not with min_score, i.e. score == min_score (pattern.rs:151 is strict).
int tfbs_audit_block(tfbs_ctx* ctx, tfbs_audit* out) {
    if (!ctx || !out) return TFBS_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof *out);
    if (!ctx->have_block) return fail(ctx, TFBS_ERR_STATE, "tfbs_audit_block needs a block (tfbs_submit_block / tfbs_upload_block first)");
    if (ctx->orig_patterns.empty()) return fail(ctx, TFBS_ERR_STATE, "tfbs_set_patterns has not been called");
    CK(cudaSetDevice(ctx->device));
    struct Hit {
        uint32_t region, pattern, group;
        int64_t start;
        bool operator<(const Hit& o) const {
            if (region != o.region) return region < o.region;
            if (pattern != o.pattern) return pattern < o.pattern;
            if (group != o.group) return group < o.group;
            return start < o.start;
        }
        bool operator==(const Hit& o) const { return region == o.region && pattern == o.pattern && group == o.group && start == o.start; }
    };
    auto take_hits = [&](std::vector<Hit>* v) {
        v->resize(ctx->n_matches);
        for (uint64_t i = 0; i < ctx->n_matches; ++i)
            (*v)[i] = Hit{ctx->h_m_region.as<uint32_t>()[i], ctx->h_m_pattern.as<uint32_t>()[i], ctx->h_m_group.as<uint32_t>()[i],
                          ctx->h_m_start.as<int64_t>()[i]};
        std::sort(v->begin(), v->end());
    };
    const int keep_record = ctx->record_matches;
    ctx->record_matches = 1;  // forces the full scan of every distinct haplotype
    ctx->audit = 1;
    std::vector<tfbs_pattern> lowered = ctx->orig_patterns;
    for (tfbs_pattern& p : lowered)
        if (p.min_score > INT32_MIN) p.min_score -= 1;  // score > INT32_MIN - 1 cannot be expressed; such a pattern has no tie list
    std::vector<Hit> with_ties, hits;
    bool overflow = false;
    const uint64_t keep_max = ctx->max_matches;
    // a run that overflows the match buffer says how many hits there were: run it once more with a buffer of that size
    auto run_recorded = [&]() -> int {
        int r = run_pipeline(ctx);
        if (r == TFBS_OK && ctx->matches_truncated && ctx->n_matches_found < (1ull << 31)) {
            ctx->max_matches = ctx->n_matches_found + ctx->n_matches_found / 8 + 1024;
            r = run_pipeline(ctx);
        }
        return r;
    };
    int rc = install_patterns(ctx, lowered.data(), (uint32_t)lowered.size());
    if (rc == TFBS_OK) rc = run_recorded();
    if (rc == TFBS_OK) {
        take_hits(&with_ties);
        overflow = ctx->matches_truncated;
    }
    // always put the caller's thresholds back, and leave the context with a normal run of the block
    int rc2 = install_patterns(ctx, ctx->orig_patterns.data(), (uint32_t)ctx->orig_patterns.size());
    if (rc2 == TFBS_OK && rc == TFBS_OK) rc2 = run_recorded();
    ctx->record_matches = keep_record;
    ctx->max_matches = keep_max;
    ctx->audit = 0;
    if (rc != TFBS_OK) return rc;
    if (rc2 != TFBS_OK) return rc2;
    take_hits(&hits);
    overflow = overflow || ctx->matches_truncated;
    ctx->tie_region.clear();
    ctx->tie_pattern.clear();
    ctx->tie_group.clear();
    ctx->tie_start.clear();
    size_t j = 0;
    for (const Hit& h : with_ties) {  // sorted set difference
        while (j < hits.size() && hits[j] < h) ++j;
        if (j < hits.size() && hits[j] == h) continue;
        ctx->tie_region.push_back(h.region);
        ctx->tie_pattern.push_back(h.pattern);
        ctx->tie_group.push_back(h.group);
        ctx->tie_start.push_back(h.start);
    }
    out->n_ties = ctx->tie_region.size();
    out->tie_region = ctx->tie_region.data();
    out->tie_pattern_index = ctx->tie_pattern.data();
    out->tie_group = ctx->tie_group.data();
    out->tie_start = ctx->tie_start.data();
    out->hap_group = ctx->h_hap_group.as<uint32_t>();
    out->hap_flags = ctx->h_hap_flags.as<uint8_t>();
    out->n_regions = ctx->R;
    out->n_samples = ctx->S;
    out->truncated = overflow ? 1 : 0;
    return TFBS_OK;
}

int tfbs_get_stats(const tfbs_ctx* ctx, tfbs_stats* out) {
    if (!ctx || !out) return TFBS_ERR_INVALID_ARGUMENT;
    *out = ctx->stats;
    return TFBS_OK;
}

int tfbs_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return TFBS_OK;
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? TFBS_OK : TFBS_ERR_CUDA;
}
int tfbs_host_unregister(void* ptr) {
    if (!ptr) return TFBS_OK;
    return cudaHostUnregister(ptr) == cudaSuccess ? TFBS_OK : TFBS_ERR_CUDA;
}

void* tfbs_stream(const tfbs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"
