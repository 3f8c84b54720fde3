#!/usr/bin/env python
"""What a per-cluster formulation of delta scoring would score, counted on the host.

Variants of a region further apart than Lmax never share a window, so the in-window variants of a region fall into clusters; the
hits of a haplotype are the reference hits plus, per cluster, a correction that depends only on WHICH of the cluster's variants it
carries.  This script counts, for a synthetic cohort, the distinct (cluster, carried subset) configurations and the window starts
they cover, next to what the current work list scores (`evaluated_cells_per_step`, `scan_items_per_step` of a bench line).
No kernel is involved; it only reads the block arrays.

    python scripts/cluster_model.py --scale 0.05
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from find_tfbs_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=2)
    args = ap.parse_args()
    pats, blk = synth.config2(scale=args.scale, seed=args.seed)
    lens = np.array([p["weights"].shape[0] for p in pats])
    lmax, sum_len = int(lens.max()), int(lens.sum())
    H = 2 * blk.n_samples
    bits = np.unpackbits(blk.carriers.view(np.uint8), axis=1, bitorder="little")[:, :H]
    n_cluster = n_config = n_pairs = starts_static = starts_dynamic = n_dyn_items = 0
    sizes = []
    for r in range(blk.n_regions):
        v = blk.variants[blk.var_off[r]:blk.var_off[r + 1]]
        keep = (v["pos"] >= blk.region_start[r]) & (v["pos"] <= blk.region_end[r])
        v = v[keep]
        if not len(v):
            continue
        order = np.argsort(v["pos"], kind="stable")
        v = v[order]
        pos, end = v["pos"], v["pos"] + v["ref_len"] - 1
        car = bits[v["carrier_row"]]                      # variants x haplotypes
        # static clusters: a variant joins the cluster while its first touched start is not past the cluster's last touched start
        cid = np.zeros(len(v), dtype=np.int64)
        cur_end = end[0]
        for i in range(1, len(v)):
            if pos[i] - lmax + 1 > cur_end:
                cid[i] = cid[i - 1] + 1
                cur_end = end[i]
            else:
                cid[i] = cid[i - 1]
                cur_end = max(cur_end, end[i])
        for c in range(cid[-1] + 1):
            m = cid == c
            sub = car[m]                                   # k x H
            k = sub.shape[0]
            sizes.append(k)
            n_cluster += 1
            masks = np.unique(sub.T, axis=0)
            masks = masks[masks.any(axis=1)]
            n_config += len(masks)
            n_pairs += int(sub.any(axis=0).sum())
            p, e = pos[m], end[m]
            alt_extra = np.maximum(v["alt_len"][m].astype(np.int64) - 1, 0)
            for mk in masks:                               # starts scored for this configuration: union of the carried variants' zones
                sel = mk.astype(bool)
                lo, hi = p[sel] - lmax + 1, e[sel] + alt_extra[sel]
                tot, ce = 0, None
                for a, b in zip(lo, hi):
                    if ce is None or a > ce:
                        tot += b - a + 1
                        ce = b
                        n_dyn_items += 1
                    elif b > ce:
                        tot += b - ce
                        ce = b
                starts_dynamic += tot
                starts_static += int(e.max() + alt_extra.max() - (p.min() - lmax + 1) + 1)
    sizes = np.array(sizes)
    print("regions %d, samples %d, patterns %d (Lmax %d, sum of lengths %d)" % (blk.n_regions, blk.n_samples, len(pats), lmax, sum_len))
    print("clusters %d (%.1f per region), variants per cluster: mean %.2f, p99 %d, max %d" %
          (n_cluster, n_cluster / blk.n_regions, sizes.mean(), np.percentile(sizes, 99), sizes.max()))
    print("(haplotype, cluster) pairs with a carried variant: %d" % n_pairs)
    print("distinct configurations: %d (%.2f per cluster)" % (n_config, n_config / n_cluster))
    print("window starts, zones of the carried variants only: %d in %d ranges -> %.3e cells" % (starts_dynamic, n_dyn_items, starts_dynamic * sum_len))
    print("window starts, whole cluster zone per configuration: %d -> %.3e cells" % (starts_static, starts_static * sum_len))
    print("scale the counts by %.0f for the full configs[1] block" % (1 / args.scale))


if __name__ == "__main__":
    main()
