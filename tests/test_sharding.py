"""N > 1 host logic on CPU: world_size-2 gloo processes shard a block by region ranges, compute their rows (the oracle stands in
for the device here; the GPU variant is in test_gpu_parity.py) and gather them on rank 0, which must equal the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from find_tfbs_b200 import sharding, synth
from find_tfbs_b200.binding import PatternSet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_region_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 100):
            costs = rng.integers(1, 100, size=n)
            rr = sharding.region_ranges(costs, world)
            assert len(rr) == world and rr[0][0] == 0 and rr[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rr, rr[1:])) and all(a <= b for a, b in rr)
            if n >= 20 * world:
                sums = [costs[a:b].sum() for a, b in rr]
                assert max(sums) < 1.5 * costs.sum() / world


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import parity_helpers as hp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pats = synth.make_pwms(4, seed=9, lmin=6, lmax=14)
    blk = synth.make_cohort(8, 30, seed=9, lmax_pattern=14, region_len=(60, 300), two_beds=True, variant_rate=0.05)
    ps = PatternSet(pats)
    shard, r0, i0 = sharding.shard_block(blk, world, rank, lmax=14)
    rows = hp.run_oracle(ps, shard, 0, False, 1)
    merged = sharding.gather_rows(rows, r0, i0)
    if rank == 0:
        full = hp.run_oracle(ps, blk, 0, False, 2)
        ok = all(np.array_equal(merged[k], full[k]) for k in ("region", "inner", "pattern_id", "vmin", "vmax", "left", "right"))
        ret.put((ok, len(full["region"])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n = ret.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert ok and n > 0
    assert all(p.exitcode == 0 for p in procs)


def _as_grouped(rows, n_regions):
    """Dense oracle rows -> grouped rows with one group per haplotype (bits = 32), the layout of tfbs_grouped_rows."""
    from find_tfbs_b200 import binding
    n, S = rows["left"].shape
    H = 2 * S
    packed = np.zeros((n, H), dtype=np.uint32)
    packed[:, 0::2], packed[:, 1::2] = rows["left"], rows["right"]
    base = packed.min(axis=1) if n else np.zeros(0, np.uint32)
    g = {"n_rows": n, "n_samples": S, "n_regions": n_regions, "region": rows["region"], "inner": rows["inner"], "pattern_id": rows["pattern_id"],
         "vmin": rows["vmin"], "vmax": rows["vmax"], "base": base, "bits": np.full(n, 32, np.uint8), "offset": np.arange(n, dtype=np.uint64) * H,
         "packed": (packed - base[:, None]).reshape(-1), "n_groups": np.full(n_regions, H, np.uint32),
         "hap_group": np.tile(np.arange(H, dtype=np.uint32), (n_regions, 1))}
    return binding.own_grouped(g)


def test_merge_sample_blocks_host():
    """tfbs_merge_sample_blocks (pure host code of the library): sample blocks in ALL_KEYS mode -> the rows counts_as_genotypes keeps
    over all samples (main.rs:450-458).  The oracle stands in for the device."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_helpers as hp
    from find_tfbs_b200 import binding
    pats = synth.make_pwms(5, seed=21, lmin=6, lmax=12, pvalue=2e-3)
    blk = synth.make_cohort(24, 12, seed=21, lmax_pattern=12, region_len=(60, 300), two_beds=True, variant_rate=0.05)
    ps = PatternSet(pats)
    full = hp.run_oracle(ps, blk, binding.ROWS_VARYING, False, 2)
    cuts = ((0, 7), (7, 16), (16, 24))
    parts = [hp.run_oracle(ps, sharding.sample_block(blk, a, b), binding.ROWS_ALL_KEYS, False, 2) for a, b in cuts]
    assert any(len(p["region"]) != len(parts[0]["region"]) for p in parts) or len(full["region"]) < len(parts[0]["region"])
    merged = binding.merge_sample_blocks([_as_grouped(p, blk.n_regions) for p in parts])
    for k in ("region", "inner", "pattern_id", "vmin", "vmax", "left", "right"):
        assert np.array_equal(merged[k], full[k]), k
    ref = sharding.merge_sample_shards(parts)
    assert np.array_equal(ref["left"], merged["left"])
    # no part at all, and parts without rows
    assert len(binding.merge_sample_blocks([])["region"]) == 0
    empty = {k: v[:0] for k, v in parts[0].items() if isinstance(v, np.ndarray)}
    assert len(binding.merge_sample_blocks([_as_grouped(empty, blk.n_regions)])["region"]) == 0
