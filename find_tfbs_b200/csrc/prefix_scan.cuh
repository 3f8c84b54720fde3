// prefix_scan.cuh -- generic exclusive scan (u32 in, u64 out)
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "dev_common.cuh"

namespace tfbs {

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

}  // namespace tfbs
