"""N > 1 on real devices: one process per GPU, region-range shards, rows gathered in rank 0's address space through the shared-memory
result arenas (no collective on the data path: regions are independent, reference main.rs:395-429), and the C++ driver with one
context per device (--devices 0,1).  With a single GPU in the box both ranks share device 0: the cross-process gather is the same."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from find_tfbs_b200 import synth
from oracle import pyoracle as ora
import file_writers as fw

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.environ.get("TFBS_B200_DRIVER") or os.path.join(ROOT, "find_tfbs_b200", "find-tfbs-b200")


def _cohort():
    pats = synth.make_pwms(10, seed=71, lmin=8, lmax=22)
    n_regions = 24 if os.environ.get("TFBS_TEST_SCALE") else 120  # TFBS_TEST_SCALE is set by the emulated run (host fibers are slow)
    blk = synth.make_cohort(40, n_regions, seed=71, lmax_pattern=22, region_len=(150, 700), two_beds=True, n_runs=3, frac_ins=0.08, frac_del=0.08)
    return pats, blk


def _rank(rank, world, tag, n_dev, q_ready, q_result, done):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from find_tfbs_b200 import binding, sharding
    import parity_helpers as hp
    pats, blk = _cohort()
    ps = binding.PatternSet(pats)
    shard, r0, i0 = sharding.shard_block(blk, world, rank, lmax=22)
    ctx = binding.Context(rank % n_dev)
    ctx.set_patterns(ps)
    arena = sharding.SharedArena(tag, rank, nbytes=8 << 20, create=True)
    ctx.set_result_arena(arena.buf)
    shard.pin()
    for _ in range(3):  # both halves of the arena are written; the third block lands in the first half again
        ctx.submit_block(shard)
        ctx.collect_grouped()
    q_ready.put((rank, r0, i0))
    if rank == 0:
        offsets = {}
        while len(offsets) < world:
            k, a, b = q_ready.get(timeout=300)
            offsets[k] = (a, b)
        parts, offs = [], []
        for k in range(world):
            a = sharding.SharedArena(tag, k)  # the other rank's rows, where its GPU put them
            g = binding.read_arena(a.buf, 0, expand=True)
            assert g is not None and g["sequence"] == 2
            parts.append({key: np.array(g[key]) for key in ("region", "inner", "pattern_id", "vmin", "vmax", "left", "right")})
            offs.append(offsets[k])
            a.close()
        merged = sharding.merge_rows(parts, offs)
        full = hp.run_oracle(ps, blk, 0, False, 4)
        ok = all(np.array_equal(merged[key], full[key]) for key in ("region", "inner", "pattern_id", "vmin", "vmax", "left", "right"))
        q_result.put((ok, len(full["region"]), [len(p["region"]) for p in parts]))
    done.wait(timeout=600)  # the arenas stay mapped until rank 0 has read them
    shard.unpin()
    ctx.set_result_arena(None)
    ctx.close()
    arena.close()


def test_two_ranks_gather_rows_through_shared_memory_arenas():
    import torch
    n_dev = max(1, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q_ready, q_result, done = ctx.Queue(), ctx.Queue(), ctx.Event()
    tag = "t%d" % os.getpid()
    world = 2
    procs = [ctx.Process(target=_rank, args=(r, world, tag, n_dev, q_ready, q_result, done)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        result = q_result.get(timeout=600)
    finally:
        done.set()
        for p in procs:
            p.join(timeout=120)
    assert result[0] and result[1] > 0 and all(n > 0 for n in result[2]), result
    assert all(p.exitcode == 0 for p in procs)


def _sample_rank(rank, world, tag, n_dev, q_ready, q_result, done):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from find_tfbs_b200 import binding, sharding
    import parity_helpers as hp
    pats, blk = synth.config4(n_regions=3, n_samples=160, seed=15, n_pwms=4)
    ps = binding.PatternSet(pats)
    cuts = ((0, 96), (96, 160))
    mine = sharding.sample_block(blk, *cuts[rank])
    ctx = binding.Context(rank % n_dev)
    ctx.set_option("rows_mode", binding.ROWS_ALL_KEYS)
    ctx.set_patterns(ps)
    arena = sharding.SharedArena(tag, rank, nbytes=8 << 20, create=True)
    ctx.set_result_arena(arena.buf)
    ctx.submit_block(mine)
    ctx.collect_grouped()
    q_ready.put(rank)
    if rank == 0:
        seen = set()
        while len(seen) < world:
            seen.add(q_ready.get(timeout=300))
        parts = []
        for k in range(world):
            a = sharding.SharedArena(tag, k)
            parts.append(binding.own_grouped(binding.read_arena(a.buf, 0)))
            a.close()
        merged = binding.merge_sample_blocks(parts)  # min != max over the samples of BOTH devices, after the gather (main.rs:450-458)
        want = sharding.merge_sample_shards([hp.run_oracle(ps, sharding.sample_block(blk, a, b), binding.ROWS_ALL_KEYS) for a, b in cuts])
        ok = all(np.array_equal(merged[key], want[key]) for key in ("region", "inner", "pattern_id", "vmin", "vmax", "left", "right"))
        q_result.put((ok, len(want["region"]), [int(p["n_rows"]) for p in parts]))
    done.wait(timeout=600)
    ctx.set_result_arena(None)
    ctx.close()
    arena.close()


def test_two_ranks_sample_blocks_merged_after_the_gather():
    """BASELINE.json configs[3]'s partition: the same regions, one SAMPLE BLOCK per device; the ALL_KEYS rows of both land in rank
    0's address space and tfbs_merge_sample_blocks applies the filter that needs every sample."""
    import torch
    n_dev = max(1, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q_ready, q_result, done = ctx.Queue(), ctx.Queue(), ctx.Event()
    tag = "s%d" % os.getpid()
    procs = [ctx.Process(target=_sample_rank, args=(r, 2, tag, n_dev, q_ready, q_result, done)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        result = q_result.get(timeout=600)
    finally:
        done.set()
        for p in procs:
            p.join(timeout=120)
    assert result[0] and result[1] > 0 and all(n >= result[1] for n in result[2][:1]), result
    assert all(p.exitcode == 0 for p in procs)


def test_driver_one_context_per_device(tmp_path):
    """--devices 0,1 (0,0 on a single-GPU box): chunks of merged regions are dealt to one worker thread + context per device, two
    blocks in flight each; the writer puts the chunks back in order: the output equals the single-device run byte for byte after
    gunzip, and the oracle's run() on the same files."""
    import torch
    n_dev = max(1, torch.cuda.device_count())
    pats, blk = _cohort()
    a = fw.cohort_to_files(blk, pats, str(tmp_path), bgzf=True, write_csi=True)
    base = [DRIVER, "--chromosome", a["chromosome"], "--input", a["bcf"], "--reference", a["reference"], "--bed", ",".join(a["beds"]),
            "--pwm_names", ",".join(a["names"]), "--pwm_file", a["pwm_file"], "--pwm_threshold_directory", a["threshold_dir"],
            "--pwm_threshold", "0.0001", "--chunk", "7", "--threads", "4"]
    outs = []
    for devs in ("0", "0,%d" % (1 % n_dev), "0,%d,0" % (1 % n_dev)):
        out = str(tmp_path / ("out_%s.vcf.gz" % devs.replace(",", "_")))
        p = subprocess.run(base + ["--output", out, "--devices", devs], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr
        outs.append(ora.gunzip_file(out))
    assert outs[0] == outs[1] == outs[2] and outs[0].count("\n") > 10
    want = ora.run(a["chromosome"], a["bcf"], a["beds"], a["reference"], None, a["pwm_file"], a["threshold_dir"], 1e-4, a["names"])

    def canon(text):  # row order and POS are not defined by the reference: sort by ID, renumber
        lines = text.strip().split("\n")
        rows = sorted(ln.split("\t", 2)[2] for ln in lines[1:])
        return [lines[0]] + rows
    assert canon(outs[0]) == canon(want)
