// k2_finish.cuh -- K2: inherited / lost reference hits, shared item counts, work counters
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "k2_worklist.cuh"

#ifndef TFBS_FINISH_LANES
#define TFBS_FINISH_LANES 0
#endif

namespace tfbs {

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
#if TFBS_FINISH_LANES
            // variant (not timed yet): the GS lanes of the sequence look up GS items at once (owner and its hit count: two dependent
            // loads per item that the loop below does one item after the other), then walk together through the items that have hits
            const u32 gshift = (threadIdx.x & 31u) & ~(GS - 1);
            const u32 gmask = ((1u << GS) - 1) << gshift;
            for (u64 w0 = i0; w0 < i1; w0 += GS) {
                const u64 w = w0 + lane;
                u32 owner = 0, n = 0;
                if (w < i1) {
                    owner = sq.items[w].owner;
                    n = sq.item_hits[owner];
                }
                u32 have = (__ballot_sync(gmask, n != 0) >> gshift) & ((1u << GS) - 1);
                while (have) {
                    const u32 l = (u32)__ffs((int)have) - 1;
                    have &= have - 1;
                    const u32 ol = __shfl_sync(gmask, owner, (int)(gshift + l));
                    const u32 nl = __shfl_sync(gmask, n, (int)(gshift + l));
                    if (lane == 0) total += nl;
                    const u32* src = sq.item_cnt + sq.item_coff[ol];
                    for (u32 k = lane; k < nkeys; k += GS) {
                        u32 v = src[k];
                        if (v) atomicAdd(&crow[k], v);
                    }
                }
            }
#else
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
#endif
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

}  // namespace tfbs
