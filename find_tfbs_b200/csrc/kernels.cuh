// kernels.cuh -- sm_100a kernels of the find-tfbs hot path.
//
//   K0  grouping      sig/group kernels     replaces load_diffs + group_by_diffs      (haplotype.rs:13-75)
//   K1  build         k_ref_prefix, k_walk  replaces patch_haplotype                  (haplotype.rs:94-156): segments + hash
//                     k_emit_list           2-bit packing (+ N mask) of the bases that are scored
//       dedup         k_seq_*               the sequence-keyed map of load_haplotypes (haplotype.rs:81-85)
//   K2  work list     k_items, k_item_*     what has to be scored (find_all_matches, main.rs:94-154; delta scoring, shared items)
//       scan          k_scan                replaces matches / apply_pwm              (pattern.rs:119-171)
//                                           + the hit -> inner-region test            (main.rs:500-505)
//       finish        k_group_finish        inherited / lost reference hits, shared item counts -> per-group count rows
//   K3  count/rows    k_rows_*              count_matches_by_sample fan-out           (main.rs:506-531)
//                                           + min/max filter of counts_as_genotypes   (main.rs:439-458)
//
// All arithmetic is integer; results are bit-exact by construction (scores are i32 sums).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tfbs.h"
#include "tables.hpp"

// Launch syntax and the dynamic shared-memory declaration are spelled as macros: tests/cuda_emu redefines them to run these very
// kernels, thread by thread, on the host of the GPU-less build container (a test of the kernel logic; not a product path).
#ifndef TFBS_LAUNCH
#define TFBS_LAUNCH(kernel, grid, block, smem, stream) kernel<<<(grid), (block), (smem), (stream)>>>
#endif
#ifndef TFBS_DYNAMIC_SHARED
#define TFBS_DYNAMIC_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace tfbs {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

// ------------------------------------------------------------------------------------------------
// Device-side views
// ------------------------------------------------------------------------------------------------

// One output segment of a patched haplotype: bases [out_start, next.out_start) come either from the
// reference window (kind 0: window index src, ref position region_start + relpos + k) or from an ALT
// allele (kind 1: allele_codes[src + k], every base at region_start + relpos, haplotype.rs:130-132).
struct Seg {
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
    const u64* ref_prefix;      // polynomial prefix hash of every window: entry ref_off[r] + r + j = sum_{t<j} val(code_t, t) * B^t
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
__device__ __forceinline__ u64 hash_val(u32 code, int rel) { return mix64(((u64)(u32)rel << 3) | code) | 1ULL; }
__device__ __forceinline__ u64 hash_pow(long long e) {  // HASH_B ^ e, negative exponents through the inverse
    u64 base = e < 0 ? HASH_BINV : HASH_B;
    u64 n = (u64)(e < 0 ? -e : e), r = 1;
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

// ------------------------------------------------------------------------------------------------
// K0: grouping of haplotypes by their Vec<Diff> (haplotype.rs:65-75)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ bool same_diff(const DevBlock& b, const tfbs_variant& x, const tfbs_variant& y) {
    if (x.pos != y.pos || x.ref_len != y.ref_len || x.alt_len != y.alt_len) return false;
    for (u32 i = 0; i < x.ref_len; ++i)
        if (b.allele_codes[x.ref_off + i] != b.allele_codes[y.ref_off + i]) return false;
    for (u32 i = 0; i < x.alt_len; ++i)
        if (b.allele_codes[x.alt_off + i] != b.allele_codes[y.alt_off + i]) return false;
    return true;
}

// One CTA per region.  Two records with equal (pos, reference, alternative) are the same Diff value
// for Vec<Diff> equality, so they share a class.
__global__ void k_variant_prep(DevBlock b, u32 r0, u32* var_class, u8* var_inwin) {
    u32 r = r0 + blockIdx.x;
    u32 v0 = b.var_off[r], v1 = b.var_off[r + 1];
    i64 s = b.region_start[r], e = b.region_end[r];
    for (u32 v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        tfbs_variant x = b.variants[v];
        var_inwin[v] = (x.pos >= s && x.pos <= e) ? 1 : 0;
        u32 cls = v - v0;
        for (u32 u = v0; u < v; ++u)
            if (same_diff(b, b.variants[u], x)) { cls = u - v0; break; }
        var_class[v] = cls;
    }
}

__device__ __forceinline__ bool carries(const DevBlock& b, u32 v, u32 h) {
    return (b.carriers[(size_t)b.variants[v].carrier_row * b.pitch + (h >> 5)] >> (h & 31)) & 1u;
}

// Thread per (region, haplotype): hash of the ordered list of carried Diff classes; 0 = no diff
// (such a haplotype stays in the reference set, main.rs:74-81,103-105).
__global__ void k_signatures(DevBlock b, u32 r0, u32 nr, u64 seed, u64* sig, u32* nd_in) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)nr * b.H) return;
    u32 r = r0 + (u32)(idx / b.H), h = (u32)(idx % b.H);
    u64 s = seed;
    u32 carried = 0, inw = 0;
    for (u32 v = b.var_off[r]; v < b.var_off[r + 1]; ++v)
        if (carries(b, v, h)) {
            s = mix64(s + b.var_class[v] + 1) * 0x9e3779b97f4a7c15ULL + carried;
            ++carried;
            inw += b.var_inwin[v];
        }
    sig[(size_t)r * b.H + h] = carried ? (mix64(s) | 1ULL) : 0ULL;
    nd_in[(size_t)r * b.H + h] = inw;
}

__device__ __forceinline__ u32 table_find_or_insert(u64* keys, u32 mask, u64 key) {
    u32 slot = (u32)(key >> 17) & mask;
    for (;;) {
        u64 prev = atomicCAS(&keys[slot], 0ULL, key);
        if (prev == 0ULL || prev == key) return slot;
        slot = (slot + 1) & mask;
    }
}
__device__ __forceinline__ u32 table_find(const u64* keys, u32 mask, u64 key) {
    u32 slot = (u32)(key >> 17) & mask;
    for (;;) {
        u64 k = keys[slot];
        if (k == key) return slot;
        if (k == 0ULL) return 0xffffffffu;
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ u64 region_key(u64 h, u32 r) { return mix64(h ^ ((u64)(r + 1) * 0xd6e8feb86659fd93ULL)) | 1ULL; }

__global__ void k_group_insert(u32 H, u32 r0, u32 nr, const u64* sig, u64* keys, u32* vals, u32 mask) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)nr * H) return;
    u32 r = r0 + (u32)(idx / H), h = (u32)(idx % H);
    u64 s = sig[(size_t)r * H + h];
    if (!s) return;
    u32 slot = table_find_or_insert(keys, mask, region_key(s, r));
    atomicMin(&vals[slot], h);
}

// leader[r,h] = smallest haplotype with the same signature; the class lists are compared exactly so
// that a hash collision is detected (and retried with another seed) instead of merging two groups.
__global__ void k_group_lookup(DevBlock b, u32 r0, u32 nr, const u64* sig, const u64* keys, const u32* vals, u32 mask, u32* leader,
                               DevStatus* st) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)nr * b.H) return;
    u32 r = r0 + (u32)(idx / b.H), h = (u32)(idx % b.H);
    u64 s = sig[(size_t)r * b.H + h];
    if (!s) { leader[(size_t)r * b.H + h] = 0xffffffffu; return; }
    u32 slot = table_find(keys, mask, region_key(s, r));
    u32 ld = vals[slot];
    leader[(size_t)r * b.H + h] = ld;
    if (ld == h) return;
    bool ok = ld < b.H;
    if (ok) {
        u32 v1 = b.var_off[r + 1];
        u32 i = b.var_off[r], j = i;
        for (;;) {
            while (i < v1 && !carries(b, i, h)) ++i;
            while (j < v1 && !carries(b, j, ld)) ++j;
            if (i == v1 || j == v1) { ok = (i == v1 && j == v1); break; }
            if (b.var_class[i] != b.var_class[j]) { ok = false; break; }
            ++i; ++j;
        }
    }
    if (!ok) atomicAdd(&st->sig_collision, 1u);
}

// One CTA per region: groups are numbered 1.. in order of their smallest haplotype; 0 is the reference.
__global__ void k_group_rank(u32 H, u32 r0, const u32* leader, const u32* nd_in, u32* hap_group, u32* ngroups, u32* sum_nd) {
    __shared__ u32 s_warp[32];
    __shared__ u32 s_base;
    __shared__ u32 s_nd;
    u32 r = r0 + blockIdx.x;
    const u32* ld = leader + (size_t)r * H;
    u32* hg = hap_group + (size_t)r * H;
    if (threadIdx.x == 0) { s_base = 0; s_nd = 0; }
    __syncthreads();
    u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (u32 h0 = 0; h0 < H; h0 += blockDim.x) {
        u32 h = h0 + threadIdx.x;
        u32 flag = (h < H && ld[h] == h) ? 1u : 0u;
        u32 bal = __ballot_sync(0xffffffffu, flag);
        u32 pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        u32 woff = 0, tot = 0;
        for (u32 w = 0; w < nw; ++w) { u32 c = s_warp[w]; if (w < wid) woff += c; tot += c; }
        u32 base = s_base;
        if (flag) {
            hg[h] = 1 + base + woff + pre;
            atomicAdd(&s_nd, nd_in[(size_t)r * H + h]);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + tot;
        __syncthreads();
    }
    for (u32 h = threadIdx.x; h < H; h += blockDim.x) {
        u32 l = ld[h];
        if (l == 0xffffffffu) hg[h] = 0;
        else if (l != h) hg[h] = hg[l < H ? l : h];  // leaders were written above
    }
    if (threadIdx.x == 0) { ngroups[r] = 1 + s_base; sum_nd[r] = s_nd; }
}

// ------------------------------------------------------------------------------------------------
// Generic exclusive scan (u32 in -> u64 out), three launches; total in out[n]
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u64 block_exclusive_scan(u64 v, u64* total) {
    __shared__ u64 s_w[SCAN_THREADS / 32];
    u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u64 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (u32)o) x += y;
    }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    u64 woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) { u64 c = s_w[w]; if (w < (int)wid) woff += c; tot += c; }
    __syncthreads();
    *total = tot;
    return woff + x - v;
}

__global__ void k_prefix_tiles(const u32* in, u64 n, u64* out, u64* tile_sums) {
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
    u64 tot;
    u64 ex = block_exclusive_scan(sum, &tot);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
__global__ void k_prefix_sums(u64* tile_sums, u32 n_tiles, u64* total_out) {
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (u32 t0 = 0; t0 < n_tiles; t0 += SCAN_THREADS) {
        u32 t = t0 + threadIdx.x;
        u64 v = t < n_tiles ? tile_sums[t] : 0;
        u64 tot;
        u64 ex = block_exclusive_scan(v, &tot);
        u64 carry = s_carry;
        if (t < n_tiles) tile_sums[t] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}
__global__ void k_prefix_add(u64* out, u64 n, const u64* tile_sums) {
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u64 add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// ------------------------------------------------------------------------------------------------
// K1: haplotype build
// ------------------------------------------------------------------------------------------------

// Window starts [p0, p1] of sequence q.  With delta scoring only the windows that touch a variant are scored for a patched
// haplotype; every other window is identical (bases and positions) to a window of the reference haplotype.
struct ScanItem {
    u32 q, p0, p1;
    u32 owner;   // index of the item whose count vector this one shares (itself if it is scored)
};

// A hit of the reference haplotype, kept so that patched haplotypes can inherit or lose it.
struct RefHit {
    u32 region;
    int relpos;   // window start relative to region_start
    u32 len;      // pattern length
    u32 pid;      // pid_index of the pattern
};
struct DevRefHits {
    RefHit* buf;   // slab of `capr` entries per region of the batch
    u32* cnt;      // hits per region of the batch (may exceed capr: then the batch falls back to a full scan)
    u32 capr;
    u32 r0;        // first region of the batch
};

// Sequence table of a batch: q = gbase[r] - gbase[r0] + g.
struct DevSeqs {
    u32 n_seq;
    const u64* gbase;        // per region (block-wide index), first sequence of the region (batch-relative after -gbase0)
    u64 gbase0;
    u32* seq_region;         // [n_seq]
    u32* seq_leader;         // [n_seq] haplotype whose diff list defines the sequence (0xffffffff for the reference)
    u32* seq_nd;             // [n_seq] in-window diffs
    u64* seq_doff;           // [n_seq+1] offset into dlist
    u32* dlist;              // variant indices, sorted per sequence
    Seg* segs;               // at 2*doff + 2*q, at most 2*nd + 2 entries
    u32* seq_nseg;           // [n_seq] segments without the terminator
    u32* seq_len;            // [n_seq] bases
    u32* ent_units;          // [list] packed units of a scored list entry: bases [p0 & ~31, p1 + 64]
    u64* ent_uoff;           // [list+1] offset into pk / nm
    u64* pk;                 // 32 bases per word, 2 bits each
    u32* nm;                 // N mask, bit b = base 32u+b is N
    u64* seq_hash;           // [n_seq]
    u8* seq_flags;           // bit0 truncated, bit1 dropped (overwritten in the sequence-keyed map)
    // scan work list: ranges of window starts that have to be scored
    u32* seq_nitems;         // [n_seq]
    u64* item_off;           // [n_seq+1]
    ScanItem* items;         // flat, in sequence order
    u32 n_items_cap;
    u64* item_key;           // [items] signature of the item (delta scoring, patched haplotypes)
    u32* item_hits;          // [items] hits found in the item (owners only)
    u64* item_coff;          // [items+1] offset of the owner's count vector in item_cnt
    u32* item_cnt;           // count vectors [pid][inner] of the owners
};


__global__ void k_seq_init(u32 H, u32 r0, const u32* hap_group, const u32* leader, const u32* nd_in, DevSeqs sq) {
    u32 r = r0 + blockIdx.x;
    u64 qb = sq.gbase[r] - sq.gbase0;
    if (threadIdx.x == 0) {
        sq.seq_region[qb] = r;
        sq.seq_leader[qb] = 0xffffffffu;
        sq.seq_nd[qb] = 0;
    }
    for (u32 h = threadIdx.x; h < H; h += blockDim.x) {
        if (leader[(size_t)r * H + h] == h) {
            u64 q = qb + hap_group[(size_t)r * H + h];
            sq.seq_region[q] = r;
            sq.seq_leader[q] = h;
            sq.seq_nd[q] = nd_in[(size_t)r * H + h];
        }
    }
}

// derived Ord of Diff: (pos, reference, alternative), vectors lexicographic, A<C<G<T<N (types.rs:5-8,39-44)
__device__ __forceinline__ int cmp_codes(const u8* a, u32 na, const u8* b, u32 nb) {
    u32 n = na < nb ? na : nb;
    for (u32 i = 0; i < n; ++i)
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return na == nb ? 0 : (na < nb ? -1 : 1);
}
__device__ __forceinline__ bool diff_less(const DevBlock& b, u32 x, u32 y) {
    const tfbs_variant& vx = b.variants[x];
    const tfbs_variant& vy = b.variants[y];
    if (vx.pos != vy.pos) return vx.pos < vy.pos;
    int c = cmp_codes(b.allele_codes + vx.ref_off, vx.ref_len, b.allele_codes + vy.ref_off, vy.ref_len);
    if (c) return c < 0;
    return cmp_codes(b.allele_codes + vx.alt_off, vx.alt_len, b.allele_codes + vy.alt_off, vy.alt_len) < 0;
}

__device__ __forceinline__ void report(DevStatus* st, u64 q, i64 relpos, u32 code) {
    u64 key = (q << 32) | ((u64)((relpos + (1 << 27)) & 0xfffffff) << 4) | code;
    atomicMin(&st->err_key, key);
}

// Thread per sequence: gathers the carried in-window diffs (haplotype.rs:95), sorts them (:96) and walks
// them exactly like next_chunk (:98-153), emitting segments instead of bases.
__global__ void k_walk(DevBlock b, DevSeqs sq, DevStatus* st) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= sq.n_seq) return;
    u32 r = sq.seq_region[q];
    u32 h = sq.seq_leader[q];
    u64 doff = sq.seq_doff[q];
    u32* dl = sq.dlist + doff;
    Seg* sg = sq.segs + 2 * doff + 2 * (u64)q;
    i64 start = b.region_start[r], end = b.region_end[r];
    u64 ro = b.ref_off[r];
    i64 n_ref = (i64)(b.ref_off[r + 1] - ro);
    i64 avail_end = start + n_ref - 1;
    u32 nd = 0;
    if (h != 0xffffffffu) {
        for (u32 v = b.var_off[r]; v < b.var_off[r + 1]; ++v)
            if (b.var_inwin[v] && carries(b, v, h)) {
                // insertion sort; records come sorted by position, so this is nearly linear
                u32 k = nd++;
                while (k > 0 && diff_less(b, v, dl[k - 1])) { dl[k] = dl[k - 1]; --k; }
                dl[k] = v;
            }
    }
    u32 out = 0, ns = 0;
    bool trunc = false;
    i64 rp = start;
    auto emit_ref = [&](i64 a, i64 e2) {
        if (a < start) a = start;
        if (e2 > avail_end) e2 = avail_end;
        if (e2 >= a) {
            sg[ns++] = Seg{out, (u32)(a - start), (int)(a - start), 0u};
            out += (u32)(e2 - a + 1);
        }
    };
    u32 k = 0;
    for (;;) {
        if (k == nd) {  // haplotype.rs:100-108
            if (rp <= end) emit_ref(rp, end);
            break;
        }
        const tfbs_variant d = b.variants[dl[k]];
        if (d.pos > rp) {  // :110-114
            emit_ref(rp, d.pos - 1);
            rp = d.pos;
        } else if (d.pos == rp && d.ref_len == 1) {  // :115-135 SNV or insertion
            u8 at = (rp >= start && rp <= avail_end) ? b.ref_codes[ro + (u64)(rp - start)] : (u8)4;
            if (b.allele_codes[d.ref_off] != at) { report(st, q, rp - start, DEV_REF_MISMATCH); break; }
            sg[ns++] = Seg{out, d.alt_off, (int)(rp - start), 1u};
            out += d.alt_len;
            rp += 1;
            ++k;
        } else if (d.pos == rp && d.alt_len == 1) {  // :136-140 deletion
            sg[ns++] = Seg{out, d.alt_off, (int)(rp - start), 1u};
            out += 1;
            rp += d.ref_len;
            ++k;
        } else if (d.pos == rp) {  // :141-143
            report(st, q, rp - start, DEV_MISSING_CASE);
            break;
        } else if (rp >= end) {  // :144-146
            trunc = true;
            emit_ref(rp, rp);
            break;
        } else {  // :147-149
            trunc = true;
            break;
        }
    }
    sg[ns] = Seg{out, 0u, 0, 2u};  // terminator
    sq.seq_nseg[q] = ns;
    sq.seq_len[q] = out;
    sq.seq_flags[q] = trunc ? 1 : 0;
    {   // hash of the (nuc, pos) vector from the segments
        const u64* P = b.ref_prefix + ro + r;
        u64 hsh = 0;
        for (u32 s = 0; s < ns; ++s) {
            const u32 n = sg[s + 1].out_start - sg[s].out_start;
            if (sg[s].kind == 0) hsh += hash_pow((long long)sg[s].out_start - (long long)sg[s].src) * (P[sg[s].src + n] - P[sg[s].src]);
            else {
                u64 pw = hash_pow(sg[s].out_start);
                for (u32 x = 0; x < n; ++x) { hsh += hash_val(b.allele_codes[sg[s].src + x], sg[s].relpos) * pw; pw *= HASH_B; }
            }
        }
        sq.seq_hash[q] = hsh;
    }
    if (trunc) atomicAdd(&st->n_truncated, 1u);
}

__device__ __forceinline__ u32 seg_find(const Seg* sg, u32 ns, u32 i) {  // last segment with out_start <= i
    u32 lo = 0, hi = ns;  // sg[ns] is the terminator
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (sg[mid].out_start <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Prefix sums of val(code_t, t) * B^t over the window of every region (one CTA per region).
__global__ void k_ref_prefix(DevBlock b, u32 r0, u64* prefix) {
    __shared__ u64 s_carry;
    const u32 r = r0 + blockIdx.x;
    const u64 ro = b.ref_off[r];
    const u32 n = (u32)(b.ref_off[r + 1] - ro);
    u64* P = prefix + ro + r;
    if (threadIdx.x == 0) { s_carry = 0; P[0] = 0; }
    __syncthreads();
    for (u32 t0 = 0; t0 < n; t0 += SCAN_THREADS) {
        const u32 t = t0 + threadIdx.x;
        u64 term = t < n ? hash_val(b.ref_codes[ro + t], (int)t) * hash_pow(t) : 0ULL;
        u64 tot;
        u64 ex = block_exclusive_scan(term, &tot);
        const u64 carry = s_carry;
        if (t < n) P[t + 1] = carry + ex + term;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
}

__device__ __forceinline__ u32 seq_group(const DevSeqs& sq, u32 q) {  // group index of q inside its region
    return (u32)((u64)q + sq.gbase0 - sq.gbase[sq.seq_region[q]]);
}

__global__ void k_seq_insert(DevSeqs sq, u64* keys, u32* vals, u32 mask) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= sq.n_seq) return;
    u32 g = seq_group(sq, q);
    if (g == 0) return;  // the reference haplotype is not in the map (main.rs:129-147)
    u64 key = region_key(sq.seq_hash[q] + mix64(sq.seq_len[q]), sq.seq_region[q]);
    u32 slot = table_find_or_insert(keys, mask, key);
    atomicMin(&vals[slot], g);
}

__device__ __forceinline__ void base_at(const DevBlock& b, const DevSeqs& sq, u32 q, u32 i, u8* c, int* rel) {
    const Seg* sg = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
    u32 s = seg_find(sg, sq.seq_nseg[q], i);
    Seg cur = sg[s];
    u32 o = i - cur.out_start;
    if (cur.kind == 0) { *c = b.ref_codes[b.ref_off[sq.seq_region[q]] + cur.src + o]; *rel = cur.relpos + (int)o; }
    else { *c = b.allele_codes[cur.src + o]; *rel = cur.relpos; }
}

// A later insert with an equal key overwrites the earlier one in the reference (haplotype.rs:84); the
// winner there depends on HashMap order, here the group with the smallest first haplotype wins (same rule
// as the oracle).  The losers are dropped: their haplotypes stay in the reference set (main.rs:103-105).
__global__ void k_seq_resolve(DevBlock b, DevSeqs sq, const u64* keys, const u32* vals, u32 mask, DevStatus* st) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= sq.n_seq) return;
    u32 g = seq_group(sq, q);
    if (g == 0) return;
    u64 key = region_key(sq.seq_hash[q] + mix64(sq.seq_len[q]), sq.seq_region[q]);
    u32 slot = table_find(keys, mask, key);
    u32 w = vals[slot];
    if (w == g) return;
    u32 qw = q - g + w;
    bool same = sq.seq_len[qw] == sq.seq_len[q] && sq.seq_region[qw] == sq.seq_region[q];
    // walk the two segment lists over the union of their breakpoints: two reference-copy pieces at the same position are equal by
    // construction, anything else is compared base by base (nuc and pos)
    if (same) {
        const Seg* sa = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
        const Seg* sb = sq.segs + 2 * sq.seq_doff[qw] + 2 * (u64)qw;
        const u8* refc = b.ref_codes + b.ref_off[sq.seq_region[q]];
        const u32 len = sq.seq_len[q];
        u32 ia = 0, ib = 0, i = 0;
        while (same && i < len) {
            while (sa[ia + 1].out_start <= i) ++ia;
            while (sb[ib + 1].out_start <= i) ++ib;
            const u32 ea = sa[ia + 1].out_start, eb = sb[ib + 1].out_start;
            const u32 e = ea < eb ? ea : eb;
            const u32 da = i - sa[ia].out_start, db = i - sb[ib].out_start;
            if (sa[ia].kind == 0 && sb[ib].kind == 0) {
                same = sa[ia].relpos + (int)da == sb[ib].relpos + (int)db;
            } else {
                for (u32 x = 0; same && x < e - i; ++x) {
                    const u8 ca = sa[ia].kind == 0 ? refc[sa[ia].src + da + x] : b.allele_codes[sa[ia].src + da + x];
                    const u8 cb = sb[ib].kind == 0 ? refc[sb[ib].src + db + x] : b.allele_codes[sb[ib].src + db + x];
                    const int pa = sa[ia].relpos + (sa[ia].kind == 0 ? (int)(da + x) : 0);
                    const int pb = sb[ib].relpos + (sb[ib].kind == 0 ? (int)(db + x) : 0);
                    same = ca == cb && pa == pb;
                }
            }
            i = e;
        }
    }
    if (same) {
        sq.seq_flags[q] |= 2;
        atomicAdd(&st->n_dropped, 1u);
    } else {
        atomicAdd(&st->seq_collision, 1u);
    }
}

// hap_flags (audit only, else NULL): the flags of the haplotype's own diff list, taken before the redirect (TFBS_HAP_* bits).
__global__ void k_redirect(u32 H, u32 r0, u32 nr, DevSeqs sq, u32* hap_group, u32* ref_used, u8* hap_flags) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u64)nr * H) return;
    u32 r = r0 + (u32)(idx / H), h = (u32)(idx % H);
    u32 g = hap_group[(size_t)r * H + h];
    const u8 fl = g ? sq.seq_flags[sq.gbase[r] - sq.gbase0 + g] : (u8)0;
    if (hap_flags) hap_flags[(size_t)r * H + h] = fl;
    if (fl & 2) { g = 0; hap_group[(size_t)r * H + h] = 0; }
    if (g == 0) ref_used[r] = 1;
}

// ------------------------------------------------------------------------------------------------
// K2: PWM scan
// ------------------------------------------------------------------------------------------------

struct DevPatterns {
    const u64* table;
    const ChunkDesc* chunks;
    const RunDesc* runs;
    const int* trip_pat;
    const u32* pat_len;
    const u32* pat_pid_index;
    u32 n_chunks;
    u32 n_pid;       // distinct pattern ids
    u32 n_patterns;
    u32 max_len;
    u32 sum_len;
    u64 sum_len_sq;
};

struct DevCounts {
    u32* C;               // counts, [region][group][pid][inner]
    const u64* cbase;     // per region (block-wide), offset into C relative to cbase0
    u64 cbase0;
};

struct DevMatches {
    u32 enabled;
    u32 cap;
    u32* region;
    u32* pattern_index;
    u32* group;
    i64* start;
};

#ifndef TFBS_SCAN_WARPS
#define TFBS_SCAN_WARPS 24
#endif
#ifndef TFBS_SCAN_UNROLL
#define TFBS_SCAN_UNROLL 1   /* measured on B200: 1 -> 0.835 of the roof, 2 -> 0.775, 4 -> 0.61 (instruction cache) */
#endif
constexpr int SCAN_UNROLL = TFBS_SCAN_UNROLL;
constexpr int SCAN_WARPS = TFBS_SCAN_WARPS;          // warps per CTA; one CTA per SM shares one copy of the tables
constexpr int SCAN_CTA = SCAN_WARPS * 32;
constexpr int TILE_POS = 1024;                       // window starts staged per pass (per warp)
constexpr int PLANE_BYTES = TILE_POS / 2 + 32;       // pair codes of even / odd starts (+ halo)
constexpr int RAW_UNITS = TILE_POS / 32 + 3;
constexpr int MAX_RUNS = 16;
constexpr int MAX_PIECES = 16;                       // items (or tiles of a long item) scanned together by one warp
#ifndef TFBS_MERGE_GAP
#define TFBS_MERGE_GAP 0   /* measured on B200 (configs[1]): 0 -> 8.02 ms/step, 8 -> 8.26, 24 -> 9.05, 64 -> 12.5: short items are shared more */
#endif
constexpr int MERGE_GAP = TFBS_MERGE_GAP;            // touched ranges closer than this are scored as one item (overlapping ones always are)

// Private to one warp: a warp owns the pieces of a round, so the scan needs no CTA-wide barrier.
struct __align__(16) WarpShared {
    u64 raw_pk[RAW_UNITS];
    u32 raw_nm[RAW_UNITS];
    u8 plane[2][PLANE_BYTES];
    // pieces of this round: window starts [p0, p0 + n) of item piece_item, staged at plane position pbase; vstart = starts before
    u32 piece_p0[MAX_PIECES], piece_vstart[MAX_PIECES + 1], piece_pbase[MAX_PIECES], piece_item[MAX_PIECES];
    u32 n_pieces, pad[3];
};

struct __align__(16) CtaShared {
    RunDesc runs[MAX_RUNS];
    u32 n_runs;
    u32 pad[3];
};

template <int FIELDS>
struct HitMask;
template <>
struct HitMask<3> { static constexpr u64 value = (1ULL << 20) | (1ULL << 41) | (1ULL << 62); };
template <>
struct HitMask<2> { static constexpr u64 value = (1ULL << 31) | (1ULL << 63); };

// Everything the rare path needs, passed by pointer (the structs are __grid_constant__ kernel parameters).
struct ScanEnv {
    const DevBlock* b;
    const DevSeqs* sq;
    const DevPatterns* pt;
    const DevCounts* ct;
    const DevMatches* mt;
    const DevRefHits* rh;
    DevStatus* st;
    int delta;
};

// Rare path: a window scored above the threshold in at least one field.
__device__ __noinline__ u32 scan_on_hit(u64 hit, u32 t, u32 i, u32 item_index, ChunkDesc cd, const ScanEnv* env) {
    u32 counted = 0;  // hits of a full scan: the caller adds them to the statistics once per work grab
    const DevSeqs& sq = *env->sq;
    const DevBlock& b = *env->b;
    const DevPatterns& pt = *env->pt;
    DevStatus* st = env->st;
    const ScanItem item = sq.items[item_index];
    const u32 q = item.q;
    const u32 len = sq.seq_len[q];
    const u32 nseg = sq.seq_nseg[q];
    const Seg* sg = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
    const u32 r = sq.seq_region[q];
    const u32 g = seq_group(sq, q);
    // 0 count every hit; 1 reference haplotype under delta scoring (count + remember the hit); 2 patched haplotype under delta
    // scoring: count only windows that touch a variant, into the count vector of the (shared) item
    const u32 mode = env->delta ? (g == 0 ? 1u : 2u) : 0u;
    const u32 nk = b.inner_off[r + 1] - b.inner_off[r];
    const tfbs_inner_region* inner = b.inner + b.inner_off[r];
    const i64 region_start = b.region_start[r];
    const int bits = cd.fields == 3 ? 21 : 32;
    for (u32 f = 0; f < cd.fields; ++f) {
        if (!((hit >> (bits * f + bits - 1)) & 1ULL)) continue;
        int pi = pt.trip_pat[(size_t)(cd.trip_off + t) * 3 + f];
        if (pi < 0) continue;
        u32 L = pt.pat_len[pi];
        if (i + L > len) continue;  // pattern.rs:147-149: only complete windows
        // pos of the first base of the window (pattern.rs:156)
        u32 s = seg_find(sg, nseg, i);
        Seg cur = sg[s];
        if (mode == 2 && cur.kind == 0 && i + L <= sg[s + 1].out_start) continue;  // untouched window: inherited from the reference
        i64 hs = (i64)cur.relpos + (cur.kind == 0 ? (i64)(i - cur.out_start) : 0);
        i64 he = hs + L - 1;
        u32 pl = pt.pat_pid_index[pi];
        u32* crow;
        if (mode == 2) {
            crow = sq.item_cnt + sq.item_coff[item_index];
            atomicAdd(&sq.item_hits[item_index], 1u);
        } else {
            crow = env->ct->C + (env->ct->cbase[r] - env->ct->cbase0) + (u64)g * pt.n_pid * nk;
            if (mode == 1) {
                const u32 rr = r - env->rh->r0;
                u32 slot = atomicAdd(&env->rh->cnt[rr], 1u);
                if (slot < env->rh->capr) env->rh->buf[(u64)rr * env->rh->capr + slot] = RefHit{r, (int)hs, L, pl};
                else st->refhit_overflow = 1;
            } else {
                ++counted;
            }
        }
        for (u32 k = 0; k < nk; ++k) {
            i64 is = inner[k].start - region_start, ie = inner[k].end - region_start;
            bool ov = (hs >= is && hs <= ie) || (he >= is && he <= ie);  // inner.overlaps(match.range), range.rs:18-21
            if (ov) atomicAdd(&crow[(size_t)pl * nk + k], inner[k].multiplicity);
        }
        if (env->mt->enabled) {
            u64 slot = atomicAdd(&st->n_matches, 1ULL);
            if (slot < env->mt->cap) {
                env->mt->region[slot] = r;
                env->mt->pattern_index[slot] = (u32)pi;
                env->mt->group[slot] = g;
                env->mt->start[slot] = region_start + hs;
            }
        }
    }
    return counted;
}

// Sum of the G table words of one triple, as a balanced tree (short dependency chains).
template <int LO, int HI>
__device__ __forceinline__ u64 pair_sum(const u8* tb, const u32 (&idx)[kMaxGroups]) {
    if constexpr (HI - LO == 1) {
        return *reinterpret_cast<const u64*>(tb + LO * (kPairEntries * 8) + idx[LO]);
    } else {
        constexpr int MID = (LO + HI) / 2;
        return pair_sum<LO, MID>(tb, idx) + pair_sum<MID, HI>(tb, idx);
    }
}

// All triples of one run (same number of column pairs G): G LDS.64 + 64-bit adds per triple and lane.
template <int G, int FIELDS>
__device__ __forceinline__ void scan_run(const u8* tb, u32 n_trip, u32 t0, const u32 (&idx)[kMaxGroups], u32 i, u32 item_index,
                                         const ChunkDesc& cd, const ScanEnv* env, u32& n_counted) {
#pragma unroll SCAN_UNROLL
    for (u32 t = 0; t < n_trip; ++t) {
        u64 acc = pair_sum<0, G>(tb, idx);
        u64 hit = acc & HitMask<FIELDS>::value;
        if (hit && i != 0xffffffffu) n_counted += scan_on_hit(hit, t0 + t, i, item_index, cd, env);
        tb += G * (kPairEntries * 8);
    }
}

template <int FIELDS>
__device__ __forceinline__ void scan_dispatch(u32 G, const u8* tb, u32 n_trip, u32 t0, const u32 (&idx)[kMaxGroups], u32 i, u32 item_index,
                                              const ChunkDesc& cd, const ScanEnv* env, u32& n_counted) {
    switch (G) {
#define TFBS_CASE(N) case N: scan_run<N, FIELDS>(tb, n_trip, t0, idx, i, item_index, cd, env, n_counted); break;
        TFBS_CASE(1) TFBS_CASE(2) TFBS_CASE(3) TFBS_CASE(4) TFBS_CASE(5) TFBS_CASE(6) TFBS_CASE(7) TFBS_CASE(8)
        TFBS_CASE(9) TFBS_CASE(10) TFBS_CASE(11) TFBS_CASE(12) TFBS_CASE(13) TFBS_CASE(14) TFBS_CASE(15) TFBS_CASE(16)
#undef TFBS_CASE
    }
}

__device__ __forceinline__ u32 pair_code_bytes(u32 a, u32 b) {  // pair_entry(a, b) * 8
    u32 e = (a < 4 && b < 4) ? 4 * a + b : (a == 4 ? 16 + b : 21 + a);
    return e * 8;
}

__device__ __forceinline__ bool seq_is_scanned(const DevSeqs& sq, u32 q, const u32* ref_used) {
    if (sq.seq_flags[q] & 2) return false;                               // overwritten in the sequence-keyed map
    if (seq_group(sq, q) == 0 && !ref_used[sq.seq_region[q]]) return false;  // nobody has the reference haplotype (main.rs:129)
    return true;
}

// Hash of everything that decides the hits of an item: the segments (kind, source, position) that cover the bases
// [p0, p1 + Lmax) relative to p0, the ALT bases among them, and where the sequence ends.  Two items of one region with equal
// descriptions score identically, window by window, so one of them is scored and the other shares its count vector.
__device__ __forceinline__ u64 item_signature(const DevBlock& b, const DevSeqs& sq, u32 q, u32 p0, u32 p1, u32 max_len) {
    const Seg* sg = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
    const u32 ns = sq.seq_nseg[q];
    const u32 len = sq.seq_len[q];
    const u32 bend = p1 + max_len < len ? p1 + max_len : len;
    u64 h = mix64(((u64)(p1 - p0) << 32) ^ (bend - p0));
    for (u32 s = seg_find(sg, ns, p0); s < ns && sg[s].out_start < bend; ++s) {
        const u32 a = sg[s].out_start > p0 ? sg[s].out_start : p0;
        const u32 e = sg[s + 1].out_start < bend ? sg[s + 1].out_start : bend;
        const u32 d = a - sg[s].out_start;
        h = mix64(h ^ (((u64)(a - p0) << 40) | ((u64)sg[s].kind << 32) | (u32)(sg[s].relpos + (sg[s].kind == 0 ? (int)d : 0))));
        if (sg[s].kind == 1)
            for (u32 k = a; k < e; ++k) h = h * 0x100000001b3ULL + b.allele_codes[sg[s].src + (k - sg[s].out_start)] + 1;
    }
    return h;
}

__device__ __forceinline__ bool items_equal(const DevBlock& b, const DevSeqs& sq, const ScanItem& x, const ScanItem& y, u32 max_len) {
    if (x.p1 - x.p0 != y.p1 - y.p0 || sq.seq_region[x.q] != sq.seq_region[y.q]) return false;
    const Seg* sa = sq.segs + 2 * sq.seq_doff[x.q] + 2 * (u64)x.q;
    const Seg* sb = sq.segs + 2 * sq.seq_doff[y.q] + 2 * (u64)y.q;
    const u32 la = sq.seq_len[x.q], lb = sq.seq_len[y.q];
    const u32 ea = x.p1 + max_len < la ? x.p1 + max_len : la, eb = y.p1 + max_len < lb ? y.p1 + max_len : lb;
    if (ea - x.p0 != eb - y.p0) return false;
    u32 ia = seg_find(sa, sq.seq_nseg[x.q], x.p0), ib = seg_find(sb, sq.seq_nseg[y.q], y.p0);
    for (u32 o = 0; o < ea - x.p0;) {  // o = offset from p0
        const u32 pa = x.p0 + o, pb = y.p0 + o;
        while (sa[ia + 1].out_start <= pa) ++ia;
        while (sb[ib + 1].out_start <= pb) ++ib;
        // both must sit at the same place of the same kind of segment
        if (sa[ia].kind != sb[ib].kind) return false;
        const u32 da = pa - sa[ia].out_start, db = pb - sb[ib].out_start;
        if ((da == 0) != (db == 0) && o != 0) return false;  // a boundary in one, not in the other
        if (sa[ia].relpos + (sa[ia].kind == 0 ? (int)da : 0) != sb[ib].relpos + (sb[ib].kind == 0 ? (int)db : 0)) return false;
        const u32 na = (sa[ia + 1].out_start < ea ? sa[ia + 1].out_start : ea) - pa;
        const u32 nb = (sb[ib + 1].out_start < eb ? sb[ib + 1].out_start : eb) - pb;
        if (na != nb) return false;
        if (sa[ia].kind == 1)
            for (u32 k = 0; k < na; ++k)
                if (b.allele_codes[sa[ia].src + da + k] != b.allele_codes[sb[ib].src + db + k]) return false;
        o += na;
    }
    return true;
}

// Work list of the scan.  Without delta scoring: one item per scanned sequence, all window starts.  With delta scoring the
// reference haplotype of every region is scanned in full and a patched haplotype only where a window can differ from the
// reference: a window is untouched iff it lies inside ONE reference-copy segment (then bases and positions equal the
// reference window at the same position, so does the hit).  Touched starts: [a - Lmax + 1, e - 1] for every ALT segment
// [a, e), and [b - Lmax + 1, b - 1] around a boundary b between two reference-copy segments.
template <bool FILL>
__global__ void k_items(DevBlock b, DevSeqs sq, const u32* ref_used, u32 max_len, int delta, u64* keys, u32* vals, u32 mask) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= sq.n_seq) return;
    const u32 g = seq_group(sq, q);
    const u32 len = sq.seq_len[q];
    u32 n = 0;
    const u64 base = FILL ? sq.item_off[q] : 0;
    ScanItem* out = FILL ? sq.items + base : nullptr;
    const bool dropped = sq.seq_flags[q] & 2;
    auto put = [&](u32 a, u32 z) {
        if (FILL) {
            out[n] = ScanItem{q, a, z, (u32)(base + n)};
            if (delta && g != 0) {  // candidates for sharing: the smallest item index with this signature becomes the owner
                u64 key = region_key(item_signature(b, sq, q, a, z, max_len), sq.seq_region[q]);
                sq.item_key[base + n] = key;
                atomicMin(&vals[table_find_or_insert(keys, mask, key)], (u32)(base + n));
            }
        }
        ++n;
    };
    if (len == 0 || dropped) {
        n = 0;
    } else if (!delta || g == 0) {
        if (delta || seq_is_scanned(sq, q, ref_used)) put(0u, len - 1);
    } else {
        const Seg* sg = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
        const u32 ns = sq.seq_nseg[q];
        bool open = false;
        u32 a = 0, z = 0;
        auto add = [&](long long lo, long long hi) {  // window starts [lo, hi], ascending lo
            if (lo < 0) lo = 0;
            if (hi > (long long)len - 1) hi = (long long)len - 1;
            if (hi < lo) return;
            if (open && (u32)lo <= z + MERGE_GAP) { if ((u32)hi > z) z = (u32)hi; return; }
            if (open) put(a, z);
            a = (u32)lo; z = (u32)hi; open = true;
        };
        for (u32 s = 0; s < ns; ++s) {
            const long long bb = sg[s].out_start, e = sg[s + 1].out_start;
            if (sg[s].kind == 1) add(bb - (long long)max_len + 1, e - 1);
            else if (s > 0 && sg[s - 1].kind == 0) add(bb - (long long)max_len + 1, bb - 1);
        }
        if (open) put(a, z);
    }
    if (!FILL) sq.seq_nitems[q] = n;
}

// Decide the owner of every item (exact comparison with the candidate) and mark what has to be scored:
// score_flag[w] = 1 for reference / full items and for owners.  count_size[w] = length of the owner's count vector.
__global__ void k_item_resolve(DevBlock b, DevSeqs sq, DevPatterns pt, const u64* n_items_ptr, int delta, u32 max_len, const u64* keys,
                               const u32* vals, u32 mask, u32* score_flag, u32* count_size) {
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= *n_items_ptr) return;
    ScanItem it = sq.items[w];
    const u32 g = seq_group(sq, it.q);
    u32 owner = (u32)w;
    if (delta && g != 0) {
        u32 cand = vals[table_find(keys, mask, sq.item_key[w])];
        if (cand != (u32)w && items_equal(b, sq, it, sq.items[cand], max_len)) owner = cand;
    }
    sq.items[w].owner = owner;
    const u32 r = sq.seq_region[it.q];
    const u32 nk = b.inner_off[r + 1] - b.inner_off[r];
    score_flag[w] = owner == (u32)w ? 1u : 0u;
    count_size[w] = (delta && g != 0 && owner == (u32)w) ? pt.n_pid * nk : 0u;
    sq.item_hits[w] = 0;
}

// Compact list of the items to score: first the long ones (reference haplotypes / full scans), then the shared short ones.
__global__ void k_item_lists(DevSeqs sq, const u64* n_items_ptr, const u32* score_flag, const u64* score_idx, u32* list) {
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= *n_items_ptr || !score_flag[w]) return;
    const u64 e = score_idx[w];
    list[e] = (u32)w;
    const ScanItem it = sq.items[w];
    sq.ent_units[e] = ((it.p1 + 64) >> 5) - (it.p0 >> 5) + 1;
}

// K1, second half: pack the bases the scored entries need (2 bits per base + N mask), gathered through the segment lists.
// EMIT_LANES lanes per list entry (a short item has 4-5 units), one lane per unit of 32 bases.
constexpr u32 EMIT_LANES = 8;
__global__ void k_emit_list(DevBlock b, DevSeqs sq, const u32* list, const u64* n_list_ptr) {
    const u64 e = ((u64)blockIdx.x * blockDim.x + threadIdx.x) / EMIT_LANES;
    const u32 lane = threadIdx.x % EMIT_LANES;
    if (e >= *n_list_ptr) return;
    const ScanItem it = sq.items[list[e]];
    const u32 q = it.q;
    const u32 len = sq.seq_len[q];
    const u32 ns = sq.seq_nseg[q];
    const Seg* sg = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
    const u8* refc = b.ref_codes + b.ref_off[sq.seq_region[q]];
    const u32 ubase = it.p0 >> 5, nu = sq.ent_units[e];
    const u64 uoff = sq.ent_uoff[e];
    for (u32 u = lane; u < nu; u += EMIT_LANES) {
        const u32 i0 = (ubase + u) * 32;
        u64 pk = 0;
        u32 nm = 0;
        if (i0 < len) {
            u32 s = seg_find(sg, ns, i0);
            Seg cur = sg[s];
            u32 nxt = sg[s + 1].out_start;
            for (u32 k = 0; k < 32; ++k) {
                const u32 i = i0 + k;
                if (i >= len) break;
                while (i >= nxt) { ++s; cur = sg[s]; nxt = sg[s + 1].out_start; }
                const u32 o = i - cur.out_start;
                const u8 c = cur.kind == 0 ? refc[cur.src + o] : b.allele_codes[cur.src + o];
                pk |= (u64)(c & 3) << (2 * k);
                nm |= (c == 4 ? 1u : 0u) << k;
            }
        }
        sq.pk[uoff + u] = pk;
        sq.nm[uoff + u] = nm;
    }
}

// One launch per pattern chunk.  Persistent CTAs (one per SM) hold the chunk's tables in shared memory; every WARP
// takes `per_grab` consecutive entries of the list from an atomic counter, stages their packed bases into its private
// pair-code planes (several short items side by side, long items in tiles) and scans all triples of the chunk, 32 window
// starts at a time.
template <int FIELDS>
__global__ void __launch_bounds__(SCAN_CTA, 1)
    k_scan(const __grid_constant__ DevBlock b, const __grid_constant__ DevSeqs sq, const __grid_constant__ DevPatterns pt,
           const __grid_constant__ DevCounts ct, const __grid_constant__ DevMatches mt, const __grid_constant__ DevRefHits rh,
           const u32* list, const u64* n_list_ptr, u32 per_grab, DevStatus* st, u32 chunk, int delta) {
    TFBS_DYNAMIC_SHARED(smem_raw);
    CtaShared* cs = reinterpret_cast<CtaShared*>(smem_raw);
    WarpShared* ws = reinterpret_cast<WarpShared*>(smem_raw + sizeof(CtaShared)) + (threadIdx.x >> 5);
    u8* tbl = smem_raw + sizeof(CtaShared) + SCAN_WARPS * sizeof(WarpShared);
    const u32 tid = threadIdx.x, lane = tid & 31;
    const ChunkDesc cd = pt.chunks[chunk];
    {  // tables: 128-bit coalesced copies (chunks are 16-byte aligned and padded)
        const uint4* src = reinterpret_cast<const uint4*>(pt.table + cd.tbl_off);
        uint4* dst = reinterpret_cast<uint4*>(tbl);
        u32 n16 = (cd.tbl_words + 1) / 2;
        for (u32 k = tid; k < n16; k += SCAN_CTA) dst[k] = src[k];
        if (tid < cd.n_runs && tid < MAX_RUNS) cs->runs[tid] = pt.runs[cd.run_off + tid];
        if (tid == 0) cs->n_runs = cd.n_runs;
    }
    __syncthreads();
    const u32 n_runs = cs->n_runs;
    const u64 n_list = *n_list_ptr;
    const ScanEnv env{&b, &sq, &pt, &ct, &mt, &rh, st, delta};

    for (;;) {
        u32 w = 0;
        if (lane == 0) w = atomicAdd(&st->work_counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        u64 li = (u64)w * per_grab;
        if (li >= n_list) break;
        const u64 lend = li + per_grab < n_list ? li + per_grab : n_list;
        u32 done_in_item = 0;  // starts of entry li already scored
        u32 n_counted = 0;
        while (li < lend) {
            __syncwarp();
            // a round: pack entries (long items in tiles of TILE_POS starts) into the planes until they are full
            u32 np = 0, pos_used = 0, vtot = 0;
            while (li < lend && np < MAX_PIECES) {
                const u32 item_index = list[li];
                const ScanItem item = sq.items[item_index];
                const u32 q = item.q;
                const u32 len = sq.seq_len[q];
                const u32 left = item.p1 - item.p0 + 1 - done_in_item;
                const u32 n = left < (u32)TILE_POS ? left : (u32)TILE_POS;
                const u32 blk = (n + 32 + 31) & ~31u;
                if (pos_used + blk > 2 * (PLANE_BYTES - 16) && np > 0) break;
                const u32 p0 = item.p0 + done_in_item;
                if (lane == 0) { ws->piece_p0[np] = p0; ws->piece_vstart[np] = vtot; ws->piece_pbase[np] = pos_used; ws->piece_item[np] = item_index; }
                {   // stage this piece: packed bases of [p0, p0 + blk + 1) -> pair codes at plane positions pos_used ..
                    const u64* gpk = sq.pk + sq.ent_uoff[li];
                    const u32* gnm = sq.nm + sq.ent_uoff[li];
                    const u32 n_units = sq.ent_units[li];
                    const u32 u0 = p0 / 32 - item.p0 / 32, o = p0 & 31;
                    const u32 nu = (o + blk + 1) / 32 + 1;
                    __syncwarp();
                    for (u32 k = lane; k < nu; k += 32) {
                        bool in = u0 + k < n_units;
                        ws->raw_pk[k] = in ? gpk[u0 + k] : 0ULL;
                        ws->raw_nm[k] = in ? gnm[u0 + k] : 0xffffffffu;
                    }
                    __syncwarp();
                    for (u32 j = lane; j < blk; j += 32) {
                        u32 x0 = o + j, x1 = x0 + 1;
                        u32 a = (u32)(ws->raw_pk[x0 >> 5] >> (2 * (x0 & 31))) & 3u;
                        u32 bb = (u32)(ws->raw_pk[x1 >> 5] >> (2 * (x1 & 31))) & 3u;
                        if (((ws->raw_nm[x0 >> 5] >> (x0 & 31)) & 1u) || p0 + j >= len) a = 4;
                        if (((ws->raw_nm[x1 >> 5] >> (x1 & 31)) & 1u) || p0 + j + 1 >= len) bb = 4;
                        const u32 jj = pos_used + j;
                        ws->plane[jj & 1][jj >> 1] = (u8)pair_code_bytes(a, bb);
                    }
                }
                pos_used += blk;
                vtot += n;
                ++np;
                done_in_item += n;
                if (done_in_item == item.p1 - item.p0 + 1) { ++li; done_in_item = 0; }
            }
            if (lane == 0) { ws->piece_vstart[np] = vtot; ws->n_pieces = np; }
            __syncwarp();
            for (u32 v0 = 0; v0 < vtot; v0 += 32) {
                const u32 v = v0 + lane;
                u32 k = 0;
                while (k + 1 < np && v >= ws->piece_vstart[k + 1]) ++k;
                const bool valid = v < vtot;
                const u32 off = valid ? v - ws->piece_vstart[k] : 0u;
                const u32 j = ws->piece_pbase[k] + off;
                const u32 i = valid ? ws->piece_p0[k] + off : 0xffffffffu;
                const u32 item_index = ws->piece_item[k];
                const u8* pl = &ws->plane[j & 1][j >> 1];
                u32 idx[kMaxGroups];
#pragma unroll
                for (int gg = 0; gg < kMaxGroups; ++gg) idx[gg] = pl[gg];
                const u8* tb = tbl;
                u32 t0 = 0;
                for (u32 rn = 0; rn < n_runs; ++rn) {
                    const RunDesc rd = cs->runs[rn];
                    scan_dispatch<FIELDS>(rd.groups, tb, rd.n_triples, t0, idx, i, item_index, cd, &env, n_counted);
                    tb += (size_t)rd.n_triples * rd.groups * (kPairEntries * 8);
                    t0 += rd.n_triples;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_counted += __shfl_xor_sync(0xffffffffu, n_counted, o);
        if (lane == 0 && n_counted) atomicAdd(&st->n_hits, (u64)n_counted);
    }
}

// Delta scoring, second half, one warp per sequence:
//  * a hit of the region's reference haplotype is inherited by a patched haplotype iff the hit's window lies inside ONE of its
//    reference-copy segments; otherwise it is taken back from the haplotype's count row (the rows hold differences to the
//    reference row, in wrapping u32 arithmetic);
//  * the count vectors of the (shared) items the haplotype is made of are added to its row.
__global__ void k_group_finish(DevBlock b, DevSeqs sq, DevPatterns pt, DevCounts ct, DevRefHits rh, const u32* ref_used, DevStatus* st) {
    __shared__ unsigned long long s_hits;  // one global atomic per CTA: a single address takes ~1 atomic per clock
    if (threadIdx.x == 0) s_hits = 0;
    __syncthreads();
    // 8 lanes per sequence: the work per sequence is a handful of dependent loads, so more sequences in flight hide the latency
    constexpr u32 GS = 8;
    const u32 q = (blockIdx.x * blockDim.x + threadIdx.x) / GS;
    const u32 lane = threadIdx.x % GS;
    u32 total = 0;
    if (q < sq.n_seq) {
        const u32 g = seq_group(sq, q);
        const u32 r = sq.seq_region[q];
        const u32 nh = min(rh.cnt[r - rh.r0], rh.capr);
        if (g == 0) {  // the reference haplotype keeps all of its hits, if anybody has it (main.rs:129)
            if (lane == 0 && ref_used[r]) total = nh;
        } else if (!(sq.seq_flags[q] & 2)) {  // not overwritten in the sequence-keyed map
            const u32 nk = b.inner_off[r + 1] - b.inner_off[r];
            const u32 nkeys = pt.n_pid * nk;
            const tfbs_inner_region* inner = b.inner + b.inner_off[r];
            const i64 rs = b.region_start[r];
            u32* crow = ct.C + (ct.cbase[r] - ct.cbase0) + (u64)g * nkeys;
            const Seg* sg = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
            const u32 ns = sq.seq_nseg[q];
            const RefHit* hits = rh.buf + (u64)(r - rh.r0) * rh.capr;
            for (u32 j = lane; j < nh; j += GS) {
                const RefHit h = hits[j];
                bool inside = false;
                for (u32 s = 0; s < ns && !inside; ++s)
                    inside = sg[s].kind == 0 && sg[s].relpos <= h.relpos &&
                             (i64)h.relpos + h.len <= (i64)sg[s].relpos + (i64)(sg[s + 1].out_start - sg[s].out_start);
                if (inside) { ++total; continue; }
                const i64 hs = h.relpos, he = hs + h.len - 1;
                for (u32 k = 0; k < nk; ++k) {
                    i64 is = inner[k].start - rs, ie = inner[k].end - rs;
                    if ((hs >= is && hs <= ie) || (he >= is && he <= ie)) atomicSub(&crow[(u64)h.pid * nk + k], inner[k].multiplicity);
                }
            }
            const u64 i0 = sq.item_off[q], i1 = sq.item_off[q + 1];
            for (u64 w = i0; w < i1; ++w) {
                const u32 owner = sq.items[w].owner;
                const u32 n = sq.item_hits[owner];
                if (!n) continue;
                if (lane == 0) total += n;
                const u32* src = sq.item_cnt + sq.item_coff[owner];
                for (u32 k = lane; k < nkeys; k += GS) {
                    u32 v = src[k];
                    if (v) atomicAdd(&crow[k], v);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(&s_hits, (unsigned long long)total);
    __syncthreads();
    if (threadIdx.x == 0 && s_hits) atomicAdd(&st->n_hits, s_hits);
}

// executed cells = sum over scanned sequences and patterns of max(0, len - L + 1) * L (pattern.rs:147-150)
__global__ void k_seq_stats(DevSeqs sq, DevPatterns pt, const u32* ref_used, DevStatus* st) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    u64 cells = 0;
    u32 scanned = 0;
    if (q < sq.n_seq && seq_is_scanned(sq, q, ref_used)) {
        scanned = 1;
        u32 len = sq.seq_len[q];
        if (len >= pt.max_len) cells = (u64)(len + 1) * pt.sum_len - (pt.sum_len_sq + pt.sum_len);
        else
            for (u32 p = 0; p < pt.n_patterns; ++p) {
                u32 L = pt.pat_len[p];
                if (L && len >= L) cells += (u64)(len - L + 1) * L;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cells += __shfl_xor_sync(0xffffffffu, cells, o);
        scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
    }
    if ((threadIdx.x & 31) == 0 && scanned) {
        atomicAdd(&st->executed_cells, cells);
        atomicAdd(&st->n_scanned, (u64)scanned);
    }
}

// evaluated cells = what k_scan really scored: complete windows starting inside the scored items
__global__ void k_item_stats(DevSeqs sq, DevPatterns pt, const u32* list, const u64* n_list_ptr, DevStatus* st) {
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 cells = 0;
    if (w < *n_list_ptr) {
        const ScanItem it = sq.items[list[w]];
        const u32 len = sq.seq_len[it.q];
        if (it.p1 + pt.max_len <= len) cells = (u64)(it.p1 - it.p0 + 1) * pt.sum_len;
        else
            for (u32 p = 0; p < pt.n_patterns; ++p) {
                u32 L = pt.pat_len[p];
                if (!L || len < L) continue;
                u32 last = len - L < it.p1 ? len - L : it.p1;
                if (last >= it.p0) cells += (u64)(last - it.p0 + 1) * L;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(&st->evaluated_cells, cells);
}

// ------------------------------------------------------------------------------------------------
// K3: fan-out to samples, min/max filter, row compaction
// ------------------------------------------------------------------------------------------------

// One CTA per region, one thread per key (pid, inner): v[s] = C[group(left)] + C[group(right)]
// (main.rs:441-448), min and max over samples (:450-451).  flag: 1 = row is emitted.
__global__ void k_rows_minmax(DevBlock b, u32 r0, const u32* hap_group, DevCounts ct, const u64* gbase, u32 n_pid, const u64* kbase,
                              u64 kbase0, int rows_mode, int delta, u32* vmin, u32* vmax, u32* flag, u32* max_count) {
    u32 r = r0 + blockIdx.x;
    u32 row_max = 0;
    u32 nk = b.inner_off[r + 1] - b.inner_off[r];
    u32 nkeys = n_pid * nk;
    const u32* C = ct.C + (ct.cbase[r] - ct.cbase0);
    const u32* hg = hap_group + (size_t)r * b.H;
    u64 ko = kbase[r] - kbase0;
    (void)gbase;
    for (u32 key = threadIdx.x; key < nkeys; key += blockDim.x) {
        u32 lo = 0xffffffffu, hi = 0;
        // under delta scoring the rows of patched haplotypes hold differences to the reference row (wrapping u32)
        const u32 base = delta ? C[key] : 0u;
        for (u32 s = 0; s < b.S; ++s) {
            u32 g0 = hg[2 * s], g1 = hg[2 * s + 1];
            u32 v = C[(size_t)g0 * nkeys + key] + C[(size_t)g1 * nkeys + key] + (g0 ? base : 0u) + (g1 ? base : 0u);
            lo = min(lo, v);
            hi = max(hi, v);
        }
        vmin[ko + key] = lo;
        vmax[ko + key] = hi;
        // keys exist once a hit of any scanned haplotype touched the inner region (main.rs:517-528);
        // every scanned group has at least one member, so that is hi > 0
        const u32 f = (rows_mode == TFBS_ROWS_VARYING) ? (lo != hi ? 1u : 0u) : (hi > 0 ? 1u : 0u);
        flag[ko + key] = f;
        if (f) row_max = max(row_max, hi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) row_max = max(row_max, __shfl_xor_sync(0xffffffffu, row_max, o));
    if ((threadIdx.x & 31) == 0 && row_max) atomicMax(max_count, row_max);
}

struct DevRows {
    u32* region;
    u32* inner;
    u16* pattern_id;
    u32* vmin;
    u32* vmax;
    void* left;    // [rows][S] of T
    void* right;
};

// One warp per emitted row; T = u8 / u16 / u32, the narrowest type that holds every count of the batch (or u32 on request).
template <class T>
__global__ void k_rows_write(DevBlock b, u32 r0, u32 nr, const u32* hap_group, DevCounts ct, u32 n_pid, const u16* pid_list,
                             const u64* kbase, u64 kbase0, u64 n_keys, const u32* vmin, const u32* vmax, const u32* flag,
                             const u64* rowidx, DevRows rows, u64 row_base, int delta) {
    u64 key = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u32 lane = threadIdx.x & 31;
    if (key >= n_keys || !flag[key]) return;
    // region of the key: last r with kbase[r] - kbase0 <= key
    u32 lo = r0, hi = r0 + nr;
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (kbase[mid] - kbase0 <= key) lo = mid; else hi = mid;
    }
    u32 r = lo;
    u32 nk = b.inner_off[r + 1] - b.inner_off[r];
    u32 nkeys = n_pid * nk;
    u32 kk = (u32)(key - (kbase[r] - kbase0));
    u32 pidx = kk / nk, k = kk % nk;
    u64 row = row_base + rowidx[key];
    if (lane == 0) {
        rows.region[row] = r;
        rows.inner[row] = b.inner_off[r] + k;
        rows.pattern_id[row] = pid_list[pidx];
        rows.vmin[row] = vmin[key];
        rows.vmax[row] = vmax[key];
    }
    const u32* C = ct.C + (ct.cbase[r] - ct.cbase0);
    const u32* hg = hap_group + (size_t)r * b.H;
    const u32 base = delta ? C[kk] : 0u;
    T* left = reinterpret_cast<T*>(rows.left) + row * b.S;
    T* right = reinterpret_cast<T*>(rows.right) + row * b.S;
    for (u32 s = lane; s < b.S; s += 32) {
        u32 g0 = hg[2 * s], g1 = hg[2 * s + 1];
        left[s] = (T)(C[(size_t)g0 * nkeys + kk] + (g0 ? base : 0u));
        right[s] = (T)(C[(size_t)g1 * nkeys + kk] + (g1 ? base : 0u));
    }
}

// nominal cells: every haplotype of every sample scanned on its own sequence (BASELINE.md "Unit of work")
__global__ void k_nominal(DevBlock b, u32 r0, u32 nr, const u32* hap_group, DevSeqs sq, DevPatterns pt, DevStatus* st) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 cells = 0;
    if (idx < (u64)nr * b.H) {
        u32 r = r0 + (u32)(idx / b.H), h = (u32)(idx % b.H);
        u32 len = sq.seq_len[sq.gbase[r] - sq.gbase0 + hap_group[(size_t)r * b.H + h]];
        if (len >= pt.max_len) cells = (u64)(len + 1) * pt.sum_len - (pt.sum_len_sq + pt.sum_len);
        else
            for (u32 p = 0; p < pt.n_patterns; ++p) {
                u32 L = pt.pat_len[p];
                if (L && len >= L) cells += (u64)(len - L + 1) * L;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(&st->nominal_cells, cells);
}

}  // namespace tfbs
