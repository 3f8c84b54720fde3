// k3_rows.cuh -- K3 of the full scan (dense count rows per group): fan-out to samples, min/max filter, row compaction; nominal cells
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "k2_types.cuh"

namespace tfbs {

// ------------------------------------------------------------------------------------------------
// K3: fan-out to samples, min/max filter, row compaction
// ------------------------------------------------------------------------------------------------

// One CTA per region, one thread per key (pid, inner): v[s] = C[group(left)] + C[group(right)]
// (main.rs:441-448), min and max over samples (:450-451).  flag: 1 = row is emitted.
__global__ void k_rows_minmax(DevBlock b, u32 r0, const u32* hap_group, DevCounts ct, const u64* gbase, u32 n_pid, const u64* kbase,
                              u64 kbase0, int rows_mode, u32* vmin, u32* vmax, u32* flag, u32* max_count) {
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
        for (u32 s = 0; s < b.S; ++s) {
            u32 g0 = hg[2 * s], g1 = hg[2 * s + 1];
            u32 v = C[(size_t)g0 * nkeys + key] + C[(size_t)g1 * nkeys + key];
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
                             const u64* rowidx, DevRows rows, u64 row_base) {
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
    T* left = reinterpret_cast<T*>(rows.left) + row * b.S;
    T* right = reinterpret_cast<T*>(rows.right) + row * b.S;
    for (u32 s = lane; s < b.S; s += 32) {
        u32 g0 = hg[2 * s], g1 = hg[2 * s + 1];
        left[s] = (T)C[(size_t)g0 * nkeys + kk];
        right[s] = (T)C[(size_t)g1 * nkeys + kk];
    }
}

// nominal cells: every haplotype of every sample scanned on its own sequence (BASELINE.md "Unit of work")
__global__ void k_nominal(DevBlock b, u32 r0, u32 nr, const u32* hap_group, DevSeqs sq, DevPatterns pt, DevStatus* st) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 cells = 0;
    if (idx < (u64)nr * b.H && !(sq.abort && *sq.abort)) {
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
