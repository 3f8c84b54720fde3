// prefix_scan.cuh -- generic exclusive scan (u32 in, u64 out), one launch (decoupled look-back)
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "dev_common.cuh"

namespace tfbs {

// ------------------------------------------------------------------------------------------------
// Generic exclusive scan (u32 in -> u64 out); total in out[n].  The element count may live on the device (n_ptr), so that a
// pipeline whose sizes are decided by earlier kernels never has to come back to the host between two launches.
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

// Tile descriptor of the single-pass scan: the top two bits say what the low 62 bits hold.
constexpr u64 SCAN_FLAG_AGG = 1ULL << 62;   // sum of this tile only
constexpr u64 SCAN_FLAG_PRE = 2ULL << 62;   // inclusive prefix up to and including this tile
constexpr u64 SCAN_VALUE_MASK = (1ULL << 62) - 1;

// work[0] = ticket counter, work[1 + t] = descriptor of tile t; all zero before the launch.  Tiles are handed out through the
// ticket, so a tile only ever waits for tiles whose CTAs are already running (no assumption about the order CTAs are scheduled in).
// n = n_ptr ? min(*n_ptr, n_cap) : n_cap; CTAs whose tile lies behind n leave at once.
__global__ void __launch_bounds__(SCAN_THREADS) k_exclusive_scan(const u32* __restrict__ in, u64 n_cap, const u64* n_ptr, u64* out, u64* work) {
    __shared__ u32 s_tile;
    __shared__ u64 s_prefix;
    u64 n = n_cap;
    if (n_ptr) { u64 m = *n_ptr; if (m < n) n = m; }
    if (threadIdx.x == 0) s_tile = (u32)atomicAdd((unsigned long long*)&work[0], 1ULL);
    __syncthreads();
    const u32 tile = s_tile;
    const u64 base = (u64)tile * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    if ((u64)tile * SCAN_TILE >= n) {
        if (tile == 0 && threadIdx.x == 0) out[0] = 0;  // n == 0
        return;
    }
    u32 v[SCAN_ITEMS];
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
    u64 agg;
    u64 ex = block_exclusive_scan(sum, &agg);
    volatile u64* desc = work + 1;
    if (threadIdx.x == 0) {
        desc[tile] = (tile == 0 ? SCAN_FLAG_PRE : SCAN_FLAG_AGG) | agg;
        s_prefix = 0;
    }
    if (tile > 0 && threadIdx.x < 32) {  // look back, 32 predecessors at a time
        const u32 lane = threadIdx.x;
        u64 prefix = 0;
        long long t = (long long)tile - 1;
        for (;;) {
            const long long mine = t - (long long)lane;
            u64 d = 0;
            if (mine >= 0) {
                do { d = desc[mine]; } while ((d >> 62) == 0);
            } else {
                d = SCAN_FLAG_PRE;  // before the first tile: prefix 0
            }
            const u32 pre_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
            // sum the aggregates up to (and including) the nearest tile that already knows its prefix
            const u32 stop = pre_mask ? (u32)__ffs((int)pre_mask) - 1 : 32;
            u64 take = lane <= stop ? (d & SCAN_VALUE_MASK) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) take += __shfl_xor_sync(0xffffffffu, take, o);
            prefix += take;
            if (pre_mask) break;
            t -= 32;
        }
        if (lane == 0) {
            s_prefix = prefix;
            desc[tile] = SCAN_FLAG_PRE | ((prefix + agg) & SCAN_VALUE_MASK);
        }
    }
    __syncthreads();
    u64 run = s_prefix + ex;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = run;  // this thread holds the last element: the total
}

}  // namespace tfbs
