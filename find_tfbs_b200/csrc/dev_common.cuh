// dev_common.cuh -- device-side views of a block, the status word, hashing, input encoding
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "kernels_base.cuh"

namespace tfbs {

// ------------------------------------------------------------------------------------------------
// Device-side views
// ------------------------------------------------------------------------------------------------

// One output segment of a patched haplotype: bases [out_start, next.out_start) come either from the
// reference window (kind 0: window index src, ref position region_start + relpos + k) or from an ALT
// allele (kind 1: allele_codes[src + k], every base at region_start + relpos, haplotype.rs:130-132).
struct __align__(16) Seg {  // one 128-bit load / store
    u32 out_start;
    u32 src;
    int relpos;
    u32 kind;
};

struct DevBlock {
    u32 R, S, H;
    const i64* region_start;
    const i64* region_end;
    const u64* ref_off;
    const u8* ref_codes;        // 0..4 per base of the concatenated windows
    const u32* inner_off;
    const tfbs_inner_region* inner;
    const u32* var_off;
    const tfbs_variant* variants;
    const u8* allele_codes;
    const u32* carriers;
    u32 pitch;
    // derived per variant
    const u32* var_class;       // index (inside the region) of the first record with the same Diff
    const u8* var_inwin;        // region_start <= pos <= region_end (haplotype.rs:95)
    u64* var_althash;           // sum_t val(alt[t], pos - region_start) * B^t: what the record's ALT bases add to a haplotype's hash, up to B^out_start
    const u64* ref_prefix;      // polynomial prefix hash of every window: entry ref_off[r] + r + j = sum_{t<j} val(code_t, t) * B^t
    // carried-record masks (NULL = not kept): word w of haplotype h of region r sits at mask_base[r] + w * H + h (ceil(V_r / 32) words per
    // haplotype), bit v of the mask = it carries record var_off[r] + v.  Written by k_signatures; they make the exact check of a group and the gather of a
    // haplotype's diffs cost a few words instead of a walk over every record of the region.
    u32* hap_mask;
    const u32* var_row;         // carrier row of every record, contiguous (a copy of tfbs_variant::carrier_row)
    u32* var_row_out;           // the same array, written by k_variant_prep
    const u64* mask_base;
    const u32* region_dups;     // 1 = the region holds two records with the same Diff value (equal lists may then have different masks)
    u64 hash_seed;              // seed of the sequence hash (hash_val): a block whose sequence-keyed map met a hash collision is repeated with another one
};

// Error / status word: the smallest key wins so that the reported failure is deterministic.
// key = (sequence index << 32) | (relpos + 2^27) << 4 | code
enum { DEV_OK = 0, DEV_REF_MISMATCH = 1, DEV_MISSING_CASE = 2 };

struct DevStatus {
    u64 err_key;          // ~0 = none
    u64 bad_ref_base;     // first offending index in ref_bases (~0 = none)
    u64 bad_allele_base;  // same for allele_bases
    u64 n_hits;
    u64 executed_cells;
    u64 nominal_cells;
    u64 n_scanned;        // sequences scanned
    u64 n_matches;        // cursor of the match buffer
    u32 n_dropped;        // groups overwritten in the sequence-keyed map
    u32 n_truncated;      // haplotypes truncated by an overlapping variant (haplotype.rs:144-149)
    u32 sig_collision;
    u32 seq_collision;
    u32 work_counter;     // dynamic scheduler of k_scan
    u32 n_refhits;        // unused (reference hits are counted per region, DevRefHits::cnt)
    u64 evaluated_cells;  // cells the scan kernel really scored
    u32 refhit_overflow;
    u32 max_count;        // largest per-sample count of an emitted row in this batch (decides the width of the returned counts)
};


__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

// Hash of a haplotype = sum_i val(nuc_i, pos_i) * B^i (mod 2^64): the key of the map in load_haplotypes (haplotype.rs:84) is the
// (nuc, pos) vector.  A reference-copy segment contributes B^(out - src) * (P[src + n] - P[src]) with P the prefix sums over the
// region's window, so the hash of a patched haplotype costs O(segments), not O(bases).  Equal hashes are verified exactly.
constexpr u64 HASH_B = 0x9e3779b97f4a7c15ULL;      // odd => invertible mod 2^64
constexpr u64 HASH_BINV = 0xf1de83e19937733dULL;   // HASH_B * HASH_BINV == 1 (mod 2^64), checked at start-up
__device__ __forceinline__ u64 hash_val(u32 code, int rel, u64 seed) { return mix64((((u64)(u32)rel << 3) | code) ^ seed) | 1ULL; }
// HASH_B ^ e for |e| < HASH_POW_N from a table (filled once per device by k_hash_pow_init; segment offsets and indel shifts of a
// window are almost always below it), by squaring beyond; negative exponents through the inverse.
constexpr u32 HASH_POW_N = 4096;
__device__ u64 g_hash_pow[2][HASH_POW_N];
__global__ void k_hash_pow_init() {
    if (blockIdx.x == 0 && threadIdx.x < 2) {
        const u64 base = threadIdx.x ? HASH_BINV : HASH_B;
        u64 r = 1;
        for (u32 e = 0; e < HASH_POW_N; ++e) { g_hash_pow[threadIdx.x][e] = r; r *= base; }
    }
}
__device__ __forceinline__ u64 hash_pow(long long e) {
    const u64 a = (u64)(e < 0 ? -e : e);
    if (a < HASH_POW_N) return g_hash_pow[e < 0 ? 1 : 0][a];
    u64 base = e < 0 ? HASH_BINV : HASH_B, n = a, r = 1;
    while (n) {
        if (n & 1) r *= base;
        base *= base;
        n >>= 1;
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// Input encoding: ASCII -> Nucleotide code (util.rs:4-16), unknown letters are reported
// ------------------------------------------------------------------------------------------------
__global__ void k_encode(const u8* __restrict__ ascii, u8* __restrict__ codes, u64 n, u64* bad_first) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        u8 l = ascii[i], c;
        switch (l) {
            case 65: case 97: c = 0; break;
            case 67: case 99: c = 1; break;
            case 71: case 103: c = 2; break;
            case 84: case 116: c = 3; break;
            case 78: case 110: c = 4; break;
            default: c = 4; atomicMin(bad_first, i); break;
        }
        codes[i] = c;
    }
}

}  // namespace tfbs
