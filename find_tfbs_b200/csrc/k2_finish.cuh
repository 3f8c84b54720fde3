// k2_finish.cuh -- K2: work counters (executed / evaluated cells)
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "k2_worklist.cuh"

namespace tfbs {

// executed cells = sum over scanned sequences and patterns of max(0, len - L + 1) * L (pattern.rs:147-150)
__global__ void k_seq_stats(DevSeqs sq, DevPatterns pt, const u32* ref_used, DevStatus* st) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    u64 cells = 0;
    u32 scanned = 0;
    if (q < seq_count(sq) && seq_is_scanned(sq, q, ref_used)) {
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
    if (w < *n_list_ptr && !(sq.abort && *sq.abort)) {
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

}  // namespace tfbs
