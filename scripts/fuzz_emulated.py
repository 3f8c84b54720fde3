#!/usr/bin/env python
"""Randomised parity search on the emulated kernels (tests/cuda_emu): many small cohorts with nasty parameters (dense variants,
many indels, records sharing a position, N runs, tiny regions, 1-column and 32-column patterns, loose thresholds, tiny table /
scratch / reference-hit budgets) through hp.check_parity, i.e. both scan modes against the oracle, hit lists included.
Test infrastructure; needs `make -C tests/cuda_emu`.  Prints the seeds that fail.

    python scripts/fuzz_emulated.py --seconds 600 --first-seed 0
"""
import argparse
import os
import sys
import time
import traceback

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
os.environ.setdefault("TFBS_B200_LIB", os.path.join(ROOT, "tests", "cuda_emu", "libtfbs_emu.so"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from find_tfbs_b200 import binding, synth  # noqa: E402
import parity_helpers as hp  # noqa: E402


def one_case(seed):
    rng = np.random.default_rng(10_000 + seed)
    lmin = int(rng.choice([1, 2, 5, 8, 20]))
    lmax = int(min(32, lmin + rng.choice([0, 3, 12, 24])))
    pats = synth.make_pwms(int(rng.integers(1, 7)), seed=seed, lmin=lmin, lmax=lmax, pvalue=float(rng.choice([3e-2, 1e-2, 1e-3, 1e-4])),
                           both_strands=bool(rng.integers(0, 2)))
    if rng.random() < 0.2:  # an OtherPattern in the list (never matches, pattern.rs:166-168)
        pats.append({"weights": None, "min_score": 0, "pattern_id": 999, "kind": binding.PATTERN_OTHER})
    lm = max([p["weights"].shape[0] for p in pats if p.get("weights") is not None])
    lo = int(rng.choice([12, 30, 120]))
    try:
        blk = synth.make_cohort(int(rng.choice([1, 2, 7, 33, 70])), int(rng.integers(1, 10)), seed=seed, lmax_pattern=lm,
                                region_len=(lo, lo + int(rng.choice([0, 60, 400]))), gap=(int(rng.choice([1, 40, 300])), 400),
                            variant_rate=float(rng.choice([0.0, 1 / 40, 1 / 10, 1 / 3])), frac_ins=float(rng.choice([0, 0.1, 0.4])),
                            frac_del=float(rng.choice([0, 0.1, 0.4])), indel_max=int(rng.choice([1, 4, 25])), n_runs=int(rng.choice([0, 0, 3])),
                            lowercase_frac=float(rng.choice([0, 0.2])), two_beds=bool(rng.integers(0, 2)),
                            same_pos_frac=float(rng.choice([0, 0.1, 0.4])), ld_blocks=int(rng.choice([0, 0, 3])))
    except ValueError:  # the generator cannot nest a second BED set into regions this small
        return
    opts = {}
    if rng.random() < 0.3:
        opts["table_budget_kb"] = 8
    if rng.random() < 0.3:
        opts["scratch_mb"] = 64
    if rng.random() < 0.3:
        opts["refhit_cap"] = int(rng.choice([1, 8, 200]))
    if rng.random() < 0.2:
        opts["rows_width"] = 0
    if rng.random() < 0.2:
        opts["scan_format"] = 1
    if rng.random() < 0.25:
        opts["tiny_caps"] = 1  # the sync-free pipeline starts with minimal scratch: every capacity overflows once
    ps = binding.PatternSet(pats)
    mode = int(rng.integers(0, 2))
    hp.check_parity(ps, blk, rows_mode=mode, options=opts, resident=bool(rng.integers(0, 2)))
    if rng.random() < 0.4:  # grouped rows, two blocks in flight, expanded on the host
        ctx = binding.Context(0)
        try:
            ctx.set_option("rows_mode", mode)
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.set_patterns(ps)
            h = max(1, blk.n_regions // 2)
            parts = [blk.slice(0, h, compact=bool(rng.integers(0, 2))), blk.slice(h, blk.n_regions, compact=bool(rng.integers(0, 2)))]
            for b in parts:
                ctx.submit_block(b)
            for b in parts:
                hp.assert_rows_equal(ctx.collect_grouped(expand=True), hp.run_oracle(ps, b, mode, False))
        finally:
            ctx.close()
    if rng.random() < 0.3 and all(p.get("weights") is not None for p in pats):  # the audit: ties and per-haplotype flags
        oa = hp.oracle_audit(ps, blk)
        au, rows, _ = hp.gpu_audit(ps, blk)
        assert au["ties"] == oa["ties"] and np.array_equal(au["hap_flags"], oa["hap_flags"]) and not au["truncated"]
        hp.assert_rows_equal(rows, hp.run_oracle(ps, blk))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--first-seed", type=int, default=0)
    ap.add_argument("--seed", type=int, default=None, help="run this one case only (to reproduce)")
    args = ap.parse_args()
    if args.seed is not None:
        one_case(args.seed)
        print("seed %d ok" % args.seed)
        return
    t0, seed, bad = time.time(), args.first_seed, []
    while time.time() - t0 < args.seconds:
        try:
            one_case(seed)
        except (binding.TfbsError, hp.ora.OracleError) as e:
            # both sides must fail the same way; check_parity runs the oracle first, so an oracle error means the input panics
            # in the reference too (e.g. "Missing case"): not a parity failure
            if not isinstance(e, hp.ora.OracleError):
                bad.append(seed)
                print("seed %d: %s" % (seed, e), flush=True)
        except Exception:
            bad.append(seed)
            print("seed %d FAILED\n%s" % (seed, traceback.format_exc()[-1500:]), flush=True)
        seed += 1
    print("cases %d..%d: %d failures %s" % (args.first_seed, seed - 1, len(bad), bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
