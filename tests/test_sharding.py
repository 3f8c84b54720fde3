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
