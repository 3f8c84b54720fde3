// k2_worklist.cuh -- K2: the scan work list (delta scoring, shared items) and the packing of the scored bases
// Part of the sm_100a kernels of the find-tfbs hot path; included through kernels.cuh (see the map there).
#pragma once
#include "k2_types.cuh"

namespace tfbs {

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

}  // namespace tfbs
