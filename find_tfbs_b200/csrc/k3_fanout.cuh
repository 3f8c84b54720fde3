// k3_fanout.cuh -- K3, configuration path: count differences -> per-haplotype-group counts -> min/max filter -> grouped rows
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
//
// count_matches_by_sample (main.rs:500-534) gives every sample the hits of its two haplotypes; here a haplotype's count for a key
// (pattern_id, inner region) is the reference haplotype's count plus the differences of its live configurations.  One warp owns a key:
// it builds the count of every distinct haplotype (group) of the region in a shared-memory vector, then takes min and max of
// left + right over the samples (main.rs:441-451) and, if the key is kept (min != max, main.rs:456-458), writes ONE GROUPED ROW:
// the count of every group, as `bits`-wide offsets from the row's smallest count.  The region's haplotype -> group map goes to the
// host once per region, not once per row: identical count vectors are stored once.
#pragma once
#include "k2c_configs.cuh"

namespace tfbs {

struct DevFan {
    const u32* hap_group;   // [R][H]
    const u64* gbase;       // groups of region r = gbase[r + 1] - gbase[r]
    u32 n_pid;
    int rows_mode;
    u32 groups_cap;         // groups the per-warp count vector holds
    // per key
    u32* vmin;
    u32* vmax;
    u32* flag;              // 1 = the key becomes a row
    u32* k_base;            // smallest per-group count of the key's row
    u8* k_bits;             // width of a packed entry
    u64* k_off;             // first packed word of the row
    const u64* rowidx;      // exclusive scan of flag (k_row_headers)
    // rows
    u64 rows_cap, words_cap;
    u32* o_region;
    u32* o_inner;
    u16* o_pid;
    u32* o_vmin;
    u32* o_vmax;
    u32* o_base;
    u8* o_bits;
    u64* o_off;
    u32* o_packed;
    const u16* pid_list;
    u32* max_count;         // largest left + right of an emitted row
};

__device__ __forceinline__ u32 bits_for(u32 span) {  // width that holds 0..span: 0, 1, 2, 4, 8, 16 or 32
    if (span == 0) return 0;
    if (span < 2) return 1;
    if (span < 4) return 2;
    if (span < 16) return 4;
    if (span < 256) return 8;
    if (span < 65536) return 16;
    return 32;
}

constexpr int FAN_THREADS = 256;
constexpr u32 FAN_KEYS = 1024;  // keys looked at per round (the list of the ones that need the count vector lives in shared memory)

// One CTA per region.  Thread per key first: a key no hit ever touched (DevConfigs::keyflag, the large majority) or whose differences
// all cancelled has the reference haplotype's count for everybody and is answered at once.  The remaining keys are taken one at a
// time by the whole CTA: the count of every group is built in ONE shared-memory vector (the warps apply different configurations at
// the same time; two configurations of different clusters can meet in a group, hence shared-memory atomics), min / max over the
// samples are reduced, and when the key becomes a row its packed counts are written at a position drawn from an atomic cursor (the
// payload order is arbitrary; k_row_headers numbers the rows in key order afterwards).
__global__ void __launch_bounds__(FAN_THREADS) k_fanout(DevBlock b, DevConfigs cf, DevFan fn) {
    TFBS_DYNAMIC_SHARED(smem_raw);
    __shared__ u32 s_n, s_red[4][FAN_THREADS / 32], s_bcast[4];
    __shared__ unsigned long long s_off;
    if (cf.plan->abort) return;
    const u32 r = blockIdx.x;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr u32 NW = FAN_THREADS / 32;
    u32* heavy = reinterpret_cast<u32*>(smem_raw);
    u32* val = heavy + FAN_KEYS;
    const u32 ng = (u32)(fn.gbase[r + 1] - fn.gbase[r]);
    if (ng > fn.groups_cap) {  // more distinct haplotypes than the shared-memory vector holds: the host repeats the run
        if (tid == 0) { atomicMax(&cf.plan->need_groups, ng); cf.plan->abort = 1; }
        return;
    }
    if (tid == 0) atomicMax(&cf.plan->need_groups, ng);
    const u32 nk = b.inner_off[r + 1] - b.inner_off[r];
    const u32 nkeys = fn.n_pid * nk;
    const u64 kb = cf.kbase[r];
    const u32 ncfg = cf.ncfg[r];
    const u64 cb = cf.cfgbase[r];
    const u32* Dr = cf.D + cf.dbase[r];
    const u32* hg = fn.hap_group + (size_t)r * b.H;
    u32 row_max = 0;
    for (u32 k0 = 0; k0 < nkeys; k0 += FAN_KEYS) {
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (u32 key = k0 + tid; key < nkeys && key < k0 + FAN_KEYS; key += FAN_THREADS) {
            bool any = false;
            if (cf.keyflag[kb + key]) {
                const u32* drow = Dr + (u64)key * ncfg;
                for (u32 c = 0; c < ncfg && !any; ++c) any = drow[c] != 0;
            }
            if (any) { heavy[atomicAdd(&s_n, 1u)] = key; continue; }
            // every haplotype has the reference haplotype's count: a row only when every key with a hit is asked for (main.rs:517-528)
            const u32 ref = cf.C0[kb + key];
            const u32 f = (fn.rows_mode == TFBS_ROWS_VARYING || ref == 0) ? 0u : 1u;
            fn.vmin[kb + key] = 2 * ref;
            fn.vmax[kb + key] = 2 * ref;
            fn.flag[kb + key] = f;
            fn.k_base[kb + key] = ref;
            fn.k_bits[kb + key] = 0;
            fn.k_off[kb + key] = 0;
            if (f) row_max = max(row_max, 2 * ref);
        }
        __syncthreads();
        const u32 n_heavy = s_n;
        for (u32 a = 0; a < n_heavy; ++a) {
            const u32 key = heavy[a];
            const u32 ref = cf.C0[kb + key];
            const u32* drow = Dr + (u64)key * ncfg;
            for (u32 g = tid; g < ng; g += FAN_THREADS) val[g] = 0;
            __syncthreads();
            for (u32 c = wid; c < ncfg; c += NW) {  // a warp per configuration, lanes over its members
                const u32 d = drow[c];
                if (!d) continue;
                const u64 m0 = cf.moff[cb + c], m1 = cf.moff[cb + c + 1];
                for (u64 m = m0 + lane; m < m1; m += 32) atomicAdd(&val[cf.members[m]], d);
            }
            __syncthreads();
            // smallest / largest count over the groups (the packing base and width) and min / max of left + right over the samples
            // (main.rs:441-451)
            u32 gmin = ref, gmax = ref, lo = 0xffffffffu, hi = 0;
            for (u32 g = tid; g < ng; g += FAN_THREADS) {
                const u32 c = ref + val[g];
                gmin = min(gmin, c);
                gmax = max(gmax, c);
            }
            for (u32 s = tid; s < b.S; s += FAN_THREADS) {
                const u32 v = 2 * ref + val[hg[2 * s]] + val[hg[2 * s + 1]];
                lo = min(lo, v);
                hi = max(hi, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                gmin = min(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
                gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
            }
            if (lane == 0) { s_red[0][wid] = lo; s_red[1][wid] = hi; s_red[2][wid] = gmin; s_red[3][wid] = gmax; }
            __syncthreads();
            if (tid == 0) {
                for (u32 w = 1; w < NW; ++w) {
                    lo = min(lo, s_red[0][w]);
                    hi = max(hi, s_red[1][w]);
                    gmin = min(gmin, s_red[2][w]);
                    gmax = max(gmax, s_red[3][w]);
                }
                // keys exist once a hit of any scanned haplotype touched the inner region (main.rs:517-528): hi > 0
                const u32 f = (fn.rows_mode == TFBS_ROWS_VARYING) ? (lo != hi ? 1u : 0u) : (hi > 0 ? 1u : 0u);
                const u32 bits = bits_for(gmax - gmin);
                const u32 words = (u32)(((u64)ng * bits + 31) / 32);
                unsigned long long off = 0;
                if (f) off = atomicAdd((unsigned long long*)&cf.plan->rowwords_alloc, (unsigned long long)words);
                fn.vmin[kb + key] = lo;
                fn.vmax[kb + key] = hi;
                fn.flag[kb + key] = f;
                fn.k_base[kb + key] = gmin;
                fn.k_bits[kb + key] = (u8)bits;
                fn.k_off[kb + key] = off;
                if (f) row_max = max(row_max, hi);
                s_bcast[0] = f;
                s_bcast[1] = bits;
                s_bcast[2] = gmin;
                s_bcast[3] = words;
                s_off = off;
            }
            __syncthreads();
            const u32 bits = s_bcast[1], words = s_bcast[3];
            if (s_bcast[0] && bits && s_off + words <= fn.words_cap) {  // beyond the capacity: the gate behind this kernel raises abort
                const u32 per = 32 / bits, base = s_bcast[2];
                const u64 off = s_off;
                for (u32 w = tid; w < words; w += FAN_THREADS) {
                    u32 word = 0;
                    for (u32 x = 0; x < per; ++x) {
                        const u32 g = w * per + x;
                        if (g < ng) word |= (ref + val[g] - base) << (x * bits);
                    }
                    fn.o_packed[off + w] = word;
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) row_max = max(row_max, __shfl_xor_sync(0xffffffffu, row_max, o));
    if (lane == 0 && row_max) atomicMax(fn.max_count, row_max);
}

// Rows in key order = (region, pattern_id, inner): thread per key.
__global__ void k_row_headers(DevBlock b, DevConfigs cf, DevFan fn, const u64* n_rows_ptr) {
    if (cf.plan->abort) return;
    const u32 r = blockIdx.x;
    const u32 nk = b.inner_off[r + 1] - b.inner_off[r];
    const u32 nkeys = fn.n_pid * nk;
    const u64 kb = cf.kbase[r];
    for (u32 key = threadIdx.x; key < nkeys; key += blockDim.x) {
        if (!fn.flag[kb + key]) continue;
        const u64 row = fn.rowidx[kb + key];
        if (row >= fn.rows_cap || row >= *n_rows_ptr) continue;
        fn.o_region[row] = r;
        fn.o_inner[row] = b.inner_off[r] + key % nk;
        fn.o_pid[row] = fn.pid_list[key / nk];
        fn.o_vmin[row] = fn.vmin[kb + key];
        fn.o_vmax[row] = fn.vmax[kb + key];
        fn.o_base[row] = fn.k_base[kb + key];
        fn.o_bits[row] = fn.k_bits[kb + key];
        fn.o_off[row] = fn.k_off[kb + key];
    }
}

// Grouped rows -> the reference's (Vec<u32>, Vec<u32>) per key (main.rs:500-534), one warp per row; T = u8 / u16 / u32.
struct DevDenseRows {
    void* left;
    void* right;
};
template <class T>
__global__ void k_rows_expand(u32 S, u32 H, const u32* hap_group, const u64* n_rows_ptr, const u32* o_region, const u32* o_base,
                              const u8* o_bits, const u64* o_off, const u32* o_packed, DevDenseRows out) {
    const u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    if (row >= *n_rows_ptr) return;
    const u32* hg = hap_group + (size_t)o_region[row] * H;
    const u32 base = o_base[row], bits = o_bits[row];
    const u32* pk = o_packed + o_off[row];
    const u32 per = bits ? 32 / bits : 0, mask = bits == 32 ? 0xffffffffu : ((1u << bits) - 1);
    T* left = reinterpret_cast<T*>(out.left) + row * S;
    T* right = reinterpret_cast<T*>(out.right) + row * S;
    for (u32 s = lane; s < S; s += 32) {
        const u32 g0 = hg[2 * s], g1 = hg[2 * s + 1];
        left[s] = (T)(base + (bits ? (pk[g0 / per] >> ((g0 % per) * bits)) & mask : 0u));
        right[s] = (T)(base + (bits ? (pk[g1 / per] >> ((g1 % per) * bits)) & mask : 0u));
    }
}

// hap_group as the host gets it with grouped rows: u16 when every region has fewer than 65536 groups
__global__ void k_narrow_groups(const u32* hap_group, u64 n, u16* out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = (u16)hap_group[i];
}

}  // namespace tfbs
