// k2_worklist.cuh -- K2: the work list of the full scan and the packing of the scored bases
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "k2_types.cuh"

namespace tfbs {

__device__ __forceinline__ bool seq_is_scanned(const DevSeqs& sq, u32 q, const u32* ref_used) {
    if (sq.seq_flags[q] & 2) return false;                               // overwritten in the sequence-keyed map
    if (seq_group(sq, q) == 0 && !ref_used[sq.seq_region[q]]) return false;  // nobody has the reference haplotype (main.rs:129)
    return true;
}

// Work list of the full scan (what the reference does, main.rs:101-147): one item per scanned sequence, every window start.
template <bool FILL>
__global__ void k_full_items(DevSeqs sq, const u32* ref_used) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= seq_count(sq)) return;
    const u32 len = sq.seq_len[q];
    const u32 n = (len && seq_is_scanned(sq, q, ref_used)) ? 1u : 0u;
    if (FILL) {
        const u64 w = sq.item_off[q];
        if (n && w < sq.n_items_cap) sq.items[w] = ScanItem{q, 0u, len - 1, (u32)w};
    } else {
        sq.seq_nitems[q] = n;
    }
}

// K1, second half: pack the bases the scored entries need (2 bits per base + N mask), gathered through the segment lists.
// EMIT_LANES lanes per list entry (a short item has 4-5 units), one lane per unit of 32 bases.
constexpr u32 EMIT_LANES = 8;
__global__ void k_emit_list(DevBlock b, DevSeqs sq, const u32* list, const u64* n_list_ptr) {
    const u64 e = ((u64)blockIdx.x * blockDim.x + threadIdx.x) / EMIT_LANES;
    const u32 lane = threadIdx.x % EMIT_LANES;
    if (e >= *n_list_ptr || (sq.abort && *sq.abort)) return;
    const ScanItem it = sq.items[list[e]];
    const u32 q = it.q;
    const u32 len = sq.seq_len[q];
    const u32 ns = sq.seq_nseg[q];
    const Seg* sg = sq.segs + 2 * sq.seq_doff[q] + 2 * (u64)q;
    const u8* refc = b.ref_codes + b.ref_off[sq.seq_region[q]];
    const u32 ubase = it.p0 >> 5, nu = sq.ent_units[e];
    const u64 uoff = sq.ent_uoff[e];
    if (sq.units_cap && uoff + nu > sq.units_cap) return;  // the packed bases do not fit: the run is repeated (need_units)
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

}  // namespace tfbs
