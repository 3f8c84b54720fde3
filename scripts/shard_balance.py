#!/usr/bin/env python
"""Balance of the region-range shards of the configs[2] block: every shard of a world of N run alone on one GPU (resident, dual
stream), its step time next to its cost estimate.  The N-GPU step is the slowest shard: max / mean is what imbalance costs.
usage: python scripts/shard_balance.py [N ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from find_tfbs_b200 import binding, sharding, synth


def main():
    worlds = [int(a) for a in sys.argv[1:]] or [8]
    pats, blk = synth.config3(scale=1.0, seed=3)
    lmax = max(p["weights"].shape[0] for p in pats)
    ps = binding.PatternSet(pats)
    costs = sharding.block_costs(blk, lmax=lmax)
    out = {}
    for world in worlds:
        ms, est, regs = [], [], []
        for rank in range(world):
            shard, r0, i0 = sharding.shard_block(blk, world, rank, lmax=lmax, compact=True)
            ctx = binding.Context(0)
            ctx.set_option("rows_width", 0)
            ctx.set_option("dual_stream", 1)
            ctx.set_patterns(ps)
            ctx.upload_block(shard)
            stream = torch.cuda.ExternalStream(ctx.stream())

            def run(n):
                ctx.run_resident()
                for _ in range(n - 1):
                    ctx.run_resident()
                    ctx.collect_grouped()
                ctx.collect_grouped()
            run(6)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(20)
            torch.cuda.synchronize()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / 20)
            st = ctx.stats()
            est.append(float(costs[r0:r0 + shard.n_regions].sum()))
            regs.append(shard.n_regions)
            ctx.close()
        out[world] = {"ms": ms, "regions": regs, "cost_share": [e / sum(est) for e in est], "max_over_mean": max(ms) / (sum(ms) / len(ms))}
        print("world %d: max %.3f mean %.3f max/mean %.3f | ms %s | regions %s" % (world, max(ms), sum(ms) / len(ms), out[world]["max_over_mean"],
                                                                                 " ".join("%.2f" % x for x in ms), regs), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
