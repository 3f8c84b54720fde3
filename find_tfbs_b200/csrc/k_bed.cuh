// k_bed.cuh -- merged regions of all BED files (load_peak_files, bed.rs:37-45: RangeStack over the concatenated ranges, range.rs:43-87)
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
//
// RangeStack sorts the ranges by start (stable) and folds them from the left: a range joins the last merged range when that one
// `overlaps` it (range.rs:18-21), else it starts a new one.  For well-formed ranges (start <= end) and ascending starts this is:
// range i starts a new merged range iff start[i] > max(end[j], j < i) -- ends only grow inside a merged range and every earlier
// merged range ended before this one began -- so the fold is a prefix maximum, a flag and a compaction.
#pragma once
#include "prefix_scan.cuh"

namespace tfbs {

constexpr int BED_THREADS = 256;

// Stable rank by start: thread per range, the ranges pass through shared memory tile by tile (n^2 / 2 compares over the grid: BED
// sets hold 10^4 .. 10^5 ranges).  order[rank] = index.
__global__ void __launch_bounds__(BED_THREADS) k_bed_rank(const u64* __restrict__ start, u64 n, u32* __restrict__ order) {
    __shared__ u64 s_start[BED_THREADS];
    const u64 i = (u64)blockIdx.x * BED_THREADS + threadIdx.x;
    const u64 mine = i < n ? start[i] : 0;
    u32 rank = 0;
    for (u64 t0 = 0; t0 < n; t0 += BED_THREADS) {
        __syncthreads();
        s_start[threadIdx.x] = t0 + threadIdx.x < n ? start[t0 + threadIdx.x] : ~0ULL;
        __syncthreads();
        const u32 m = (u32)(n - t0 < (u64)BED_THREADS ? n - t0 : (u64)BED_THREADS);
        for (u32 k = 0; k < m; ++k) {
            const u64 s = s_start[k];
            rank += (s < mine || (s == mine && t0 + k < i)) ? 1u : 0u;
        }
    }
    if (i < n) order[rank] = (u32)i;
}

// One CTA walks the sorted ranges in tiles: prefix maximum of the ends, flag, running count of merged ranges.
// out_start / out_end [n]; *n_out = merged ranges; *bad = 1 if a range has end < start (the fold is then not a prefix maximum:
// the caller merges on the host with the literal rule).
__global__ void __launch_bounds__(BED_THREADS) k_bed_merge(const u64* __restrict__ start, const u64* __restrict__ end, const u32* __restrict__ order, u64 n,
                                                           u64* __restrict__ out_start, u64* __restrict__ out_end, u64* n_out, u32* bad) {
    __shared__ u64 s_w[BED_THREADS / 32];
    __shared__ u32 s_c[BED_THREADS / 32];
    __shared__ u64 s_carry_max;
    __shared__ u32 s_carry_cnt;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_carry_max = 0; s_carry_cnt = 0; }
    __syncthreads();
    for (u64 t0 = 0; t0 < n; t0 += BED_THREADS) {
        const u64 i = t0 + tid;
        const bool in = i < n;
        const u32 src = in ? order[i] : 0u;
        const u64 s = in ? start[src] : 0, e = in ? end[src] : 0;
        if (in && e < s) *bad = 1;
        // inclusive prefix maximum of the ends inside the tile
        u64 x = e;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (u32)o && y > x) x = y;
        }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        // exclusive maximum for this element: earlier tiles, earlier warps of the tile, earlier lanes (ends are >= 0: 0 is neutral)
        u64 ex = s_carry_max;
        for (u32 w = 0; w < wid; ++w) if (s_w[w] > ex) ex = s_w[w];
        const u64 up = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane > 0 && up > ex) ex = up;
        const u32 flag = in && (i == 0 || s > ex) ? 1u : 0u;
        // position of the merged range: carried count + flags before this element
        u32 c = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(0xffffffffu, c, o);
            if (lane >= (u32)o) c += y;
        }
        if (lane == 31) s_c[wid] = c;
        __syncthreads();
        u32 cbefore = 0, ctot = 0;
        for (u32 w = 0; w < BED_THREADS / 32; ++w) { if (w < wid) cbefore += s_c[w]; ctot += s_c[w]; }
        const u32 id = s_carry_cnt + cbefore + c - 1;  // merged range this element belongs to (valid once one has started)
        if (flag) { out_start[id] = s; out_end[id] = e; }
        __syncthreads();  // the starts of this tile are in place before the ends grow
        if (in && !flag) atomicMax((unsigned long long*)&out_end[id], (unsigned long long)e);
        __syncthreads();
        if (tid == 0) {
            u64 m = s_carry_max;
            for (u32 w = 0; w < BED_THREADS / 32; ++w) if (s_w[w] > m) m = s_w[w];
            s_carry_max = m;
            s_carry_cnt += ctot;
        }
        __syncthreads();
    }
    if (tid == 0) *n_out = s_carry_cnt;
}

}  // namespace tfbs
