// k3_fanout.cuh -- K3, configuration path: count differences -> per-haplotype-group counts -> min/max filter -> grouped rows
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
//
// count_matches_by_sample (main.rs:500-534) gives every sample the hits of its two haplotypes; here a haplotype's count for a key
// (pattern_id, inner region) is the reference haplotype's count plus the differences of its live configurations.  One warp owns a key:
// it builds the count of every distinct haplotype (group) of the region in its private shared-memory vector -- no atomics: the members
// of one configuration are distinct groups, configurations are applied one after the other -- then takes min and max of
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
    u32* flag;
    u32* rowwords;          // packed words of the key's row (0 when it is not emitted)
    const u64* rowidx;      // exclusive scans of flag / rowwords (second pass)
    const u64* rowoff;
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

constexpr int FAN_WARPS = 8;

// One CTA per region, one warp per key at a time.  WRITE = false: vmin / vmax / flag / rowwords of every key; WRITE = true: the rows.
template <bool WRITE>
__global__ void __launch_bounds__(FAN_WARPS * 32) k_fanout(DevBlock b, DevConfigs cf, DevFan fn) {
    TFBS_DYNAMIC_SHARED(smem_raw);
    if (cf.plan->abort) return;
    const u32 r = blockIdx.x;
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u32* val = reinterpret_cast<u32*>(smem_raw) + (size_t)wid * fn.groups_cap;
    const u32 ng = (u32)(fn.gbase[r + 1] - fn.gbase[r]);
    if (ng > fn.groups_cap) {  // more distinct haplotypes than the shared-memory vector holds: the host repeats the run
        if (threadIdx.x == 0) { atomicMax(&cf.plan->need_groups, ng); cf.plan->abort = 1; }
        return;
    }
    if (threadIdx.x == 0) atomicMax(&cf.plan->need_groups, ng);
    const u32 nk = b.inner_off[r + 1] - b.inner_off[r];
    const u32 nkeys = fn.n_pid * nk;
    const u64 kb = cf.kbase[r];
    const u32 ncfg = cf.ncfg[r];
    const u64 cb = cf.cfgbase[r];
    const u32* Dr = cf.D + cf.dbase[r];
    const u32* hg = fn.hap_group + (size_t)r * b.H;
    u32 row_max = 0;
    for (u32 key = wid; key < nkeys; key += nw) {
        if (WRITE && !fn.flag[kb + key]) continue;
        const u32 ref = cf.C0[kb + key];
        const u32* drow = Dr + (u64)key * ncfg;
        // which configurations change this key?
        bool any = false;
        for (u32 c0 = 0; c0 < ncfg; c0 += 32) {
            const u32 d = c0 + lane < ncfg ? drow[c0 + lane] : 0u;
            if (__ballot_sync(0xffffffffu, d != 0)) { any = true; break; }
        }
        u32 lo, hi, gmin = ref, gmax = ref;
        if (!any) {
            lo = hi = 2 * ref;  // every haplotype has the reference haplotype's count
        } else {
            for (u32 g = lane; g < ng; g += 32) val[g] = 0;
            __syncwarp();
            for (u32 c0 = 0; c0 < ncfg; c0 += 32) {
                const u32 d = c0 + lane < ncfg ? drow[c0 + lane] : 0u;
                u32 nz = __ballot_sync(0xffffffffu, d != 0);
                while (nz) {
                    const u32 l = (u32)__ffs((int)nz) - 1;
                    nz &= nz - 1;
                    const u32 dl = __shfl_sync(0xffffffffu, d, (int)l);
                    const u64 m0 = cf.moff[cb + c0 + l], m1 = cf.moff[cb + c0 + l + 1];
                    for (u64 m = m0 + lane; m < m1; m += 32) val[cf.members[m]] += dl;  // distinct groups: no conflict
                    __syncwarp();
                }
            }
            // smallest / largest count over the groups (the packing base and width) ...
            for (u32 g = lane; g < ng; g += 32) {
                const u32 c = ref + val[g];
                gmin = min(gmin, c);
                gmax = max(gmax, c);
            }
            // ... and min / max of left + right over the samples (main.rs:441-451)
            lo = 0xffffffffu;
            hi = 0;
            for (u32 s = lane; s < b.S; s += 32) {
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
        }
        const u32 bits = bits_for(gmax - gmin);
        const u32 words = (u32)(((u64)ng * bits + 31) / 32);
        if (!WRITE) {
            // keys exist once a hit of any scanned haplotype touched the inner region (main.rs:517-528): hi > 0
            const u32 f = (fn.rows_mode == TFBS_ROWS_VARYING) ? (lo != hi ? 1u : 0u) : (hi > 0 ? 1u : 0u);
            if (lane == 0) {
                fn.vmin[kb + key] = lo;
                fn.vmax[kb + key] = hi;
                fn.flag[kb + key] = f;
                fn.rowwords[kb + key] = f ? words : 0u;
            }
            if (f) row_max = max(row_max, hi);
        } else {
            const u64 row = fn.rowidx[kb + key];
            const u64 off = fn.rowoff[kb + key];
            if (row >= fn.rows_cap || off + words > fn.words_cap) continue;  // the gate in front of this pass has raised abort
            if (lane == 0) {
                fn.o_region[row] = r;
                fn.o_inner[row] = b.inner_off[r] + key % nk;
                fn.o_pid[row] = fn.pid_list[key / nk];
                fn.o_vmin[row] = lo;
                fn.o_vmax[row] = hi;
                fn.o_base[row] = gmin;
                fn.o_bits[row] = (u8)bits;
                fn.o_off[row] = off;
            }
            if (bits) {
                const u32 per = 32 / bits;
                for (u32 w = lane; w < words; w += 32) {
                    u32 word = 0;
                    for (u32 x = 0; x < per; ++x) {
                        const u32 g = w * per + x;
                        if (g < ng) word |= (ref + (any ? val[g] : 0u) - gmin) << (x * bits);
                    }
                    fn.o_packed[off + w] = word;
                }
            }
        }
        __syncwarp();
    }
    if (!WRITE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) row_max = max(row_max, __shfl_xor_sync(0xffffffffu, row_max, o));
        if (lane == 0 && row_max) atomicMax(fn.max_count, row_max);
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
