"""Multi-GPU partitioning of the hot path: merged regions are independent (reference main.rs:395-429), so the block is cut
into contiguous region ranges, one per rank, with no collective on the data path.  Rows come back in
(region, pattern_id, inner) order per rank; concatenating the ranks in order gives the single-process result."""
import numpy as np


def region_ranges(costs, world):
    """Contiguous ranges [r0, r1) per rank balancing sum(costs): costs[r] ~ window length x expected distinct haplotypes."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if n == 0:
        return [(0, 0)] * world
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    total = cum[-1]
    cuts = [0]
    for k in range(1, world):
        target = total * k / world
        r = int(np.searchsorted(cum, target, side="left"))
        # the boundary closest to the target, never before the previous cut
        if r > 0 and abs(cum[r - 1] - target) <= abs(cum[min(r, n)] - target):
            r -= 1
        cuts.append(min(n, max(cuts[-1], r)))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def block_costs(block, lmax=30):
    """Device work per region, in window starts the scan scores: the reference haplotype in full (the window length) plus about
    1.4 configurations per record (measured on the synthetic cohorts) of 2 * Lmax starts each.  The bookkeeping kernels (grouping,
    walk, fan-out) grow with the records of the region too, so this also balances them; the window length alone does not."""
    w = (block.region_end - block.region_start + 1).astype(np.float64)
    v = np.diff(block.var_off.astype(np.int64)).astype(np.float64)
    return w + 1.4 * 2.0 * lmax * v


def shard_block(block, world, rank, lmax=30, compact=True):
    """The rank's contiguous range of regions as a block of its own; returns (shard, first region, first inner index)."""
    r0, r1 = region_ranges(block_costs(block, lmax), world)[rank]
    return block.slice(r0, r1, compact=compact), r0, int(block.inner_off[r0])


def merge_rows(parts, offsets):
    """parts: per-rank row dicts with region / inner indices relative to the shard; offsets: (first region, first inner index) of
    each shard, as shard_block returns them.  Rows come back block-wide, in the order of a single-process run."""
    out = {}
    for k in ("pattern_id", "vmin", "vmax", "left", "right"):
        out[k] = np.concatenate([p[k] for p in parts]) if parts else np.zeros(0)
    for k, j in (("region", 0), ("inner", 1)):
        out[k] = (np.concatenate([p[k].astype(np.int64) + o[j] for p, o in zip(parts, offsets)]).astype(np.uint32) if parts
                  else np.zeros(0, np.uint32))
    return out


def gather_rows(rows, region_offset, inner_offset, group=None):
    """torch.distributed gather of the per-rank rows to rank 0 (host tensors; gloo or nccl object collectives).  Returns the merged
    rows on rank 0 and None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((dict(rows), (region_offset, inner_offset)), gathered, dst=0, group=group)
    if rank != 0:
        return None
    return merge_rows([g[0] for g in gathered], [g[1] for g in gathered])


class SharedArena:
    """The result arena of one rank as a file in /dev/shm: the rank's GPU copies its grouped rows straight into it
    (tfbs_set_result_arena) and the gathering rank maps the same file -- the rows of all GPUs of the box end up in one process'
    address space without a collective and without a host copy (regions are independent, reference main.rs:395-429)."""

    def __init__(self, tag, rank, nbytes=None, create=False):
        import os
        self.path = "/dev/shm/tfbs_arena_%s_%d" % (tag, rank)
        self.created = create
        if create:
            self.buf = np.memmap(self.path, dtype=np.uint8, mode="w+", shape=(int(nbytes),))
        else:
            self.buf = np.memmap(self.path, dtype=np.uint8, mode="r", shape=(os.path.getsize(self.path),))

    def close(self):
        import os
        del self.buf
        if self.created and os.path.exists(self.path):
            os.unlink(self.path)


# ---- sample-block sharding (BASELINE.json configs[3]: biobank-scale cohorts) -------------------------------------------
def sample_block(block, s0, s1):
    """The same regions and records restricted to samples [s0, s1) (carrier bits re-packed)."""
    from .binding import Block
    H = 2 * block.n_samples
    if (2 * s0) % 32 == 0:  # the cut falls on a word of the carrier rows: slice the words, clear the bits behind the last haplotype
        n_h = 2 * (s1 - s0)
        pitch = max(1, (n_h + 31) // 32)
        carriers = np.zeros((block.carriers.shape[0], pitch), dtype=np.uint32)
        src = block.carriers[:, 2 * s0 // 32:2 * s0 // 32 + pitch]
        carriers[:, :src.shape[1]] = src
        if n_h % 32:
            carriers[:, n_h // 32] &= np.uint32((1 << (n_h % 32)) - 1)
        return Block(s1 - s0, block.region_start, block.region_end, block.ref_off, block.ref_bases, block.inner_off, block.inner,
                     block.var_off, block.variants, block.allele_bases, carriers)
    bits = np.unpackbits(block.carriers.view(np.uint8), axis=1, bitorder="little")[:, :H]
    sub = bits[:, 2 * s0:2 * s1]
    pitch = max(1, (2 * (s1 - s0) + 31) // 32)
    padded = np.zeros((sub.shape[0], pitch * 32), dtype=np.uint8)
    padded[:, :sub.shape[1]] = sub
    carriers = np.packbits(padded, axis=1, bitorder="little").view(np.uint32)
    return Block(s1 - s0, block.region_start, block.region_end, block.ref_off, block.ref_bases, block.inner_off, block.inner,
                 block.var_off, block.variants, block.allele_bases, carriers)


def merge_sample_shards(parts):
    """parts: rows of every sample shard in ROWS_ALL_KEYS mode (same regions, disjoint sample ranges, in sample order).
    Counts are per sample, so shards concatenate along the sample axis; a key missing in a shard had no hit there (all zero).
    The min == max filter of counts_as_genotypes (reference main.rs:450-458) needs ALL samples, so it is applied here, after the
    gather.  Exact whenever no region has a sequence-keyed overwrite (the reference's own outcome is undefined there,
    and the deterministic winner is chosen per shard)."""
    keys = {}
    for p in parts:
        for i in range(len(p["region"])):
            keys.setdefault((int(p["region"][i]), int(p["pattern_id"][i]), int(p["inner"][i])), None)
    order = sorted(keys)
    index = {k: i for i, k in enumerate(order)}
    n = len(order)
    lefts, rights = [], []
    for p in parts:
        S = p["left"].shape[1]
        left = np.zeros((n, S), dtype=np.uint32)
        right = np.zeros((n, S), dtype=np.uint32)
        for i in range(len(p["region"])):
            j = index[(int(p["region"][i]), int(p["pattern_id"][i]), int(p["inner"][i]))]
            left[j] = p["left"][i]
            right[j] = p["right"][i]
        lefts.append(left)
        rights.append(right)
    left = np.concatenate(lefts, axis=1) if lefts else np.zeros((0, 0), np.uint32)
    right = np.concatenate(rights, axis=1) if rights else np.zeros((0, 0), np.uint32)
    v = left.astype(np.int64) + right
    vmin = v.min(axis=1) if n else np.zeros(0, np.int64)
    vmax = v.max(axis=1) if n else np.zeros(0, np.int64)
    keep = vmin != vmax
    arr = np.array(order, dtype=np.int64).reshape(-1, 3)
    return {"region": arr[keep, 0].astype(np.uint32), "pattern_id": arr[keep, 1].astype(np.uint16), "inner": arr[keep, 2].astype(np.uint32),
            "vmin": vmin[keep].astype(np.uint32), "vmax": vmax[keep].astype(np.uint32), "left": left[keep], "right": right[keep]}
