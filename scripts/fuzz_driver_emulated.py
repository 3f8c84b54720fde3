#!/usr/bin/env python
"""Randomised end-to-end search: synthetic cohorts written as BCF / FASTA / BED / PWM files (plain gzip or BGZF + CSI index, with
and without foreign contigs) through the C++ driver linked against the emulated kernels (tests/cuda_emu/find-tfbs-emu), compared
with the oracle's run() on the same files, byte for byte after gunzip.  Test infrastructure; needs `make -C tests/cuda_emu`.

    python scripts/fuzz_driver_emulated.py --seconds 600
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from find_tfbs_b200 import synth  # noqa: E402
from oracle import pyoracle as ora  # noqa: E402
import file_writers as fw  # noqa: E402

DRIVER = os.path.join(ROOT, "tests", "cuda_emu", "find-tfbs-emu")


def one_case(seed, work):
    rng = np.random.default_rng(77_000 + seed)
    pats = synth.make_pwms(int(rng.integers(1, 6)), seed=seed, lmin=int(rng.choice([4, 8])), lmax=int(rng.choice([10, 22, 30])),
                           pvalue=float(rng.choice([1e-2, 1e-3, 1e-4])))
    lm = max(p["weights"].shape[0] for p in pats)
    blk = synth.make_cohort(int(rng.choice([1, 3, 12, 40])), int(rng.integers(1, 25)), seed=seed, lmax_pattern=lm,
                            region_len=(40, int(rng.choice([80, 400]))), variant_rate=float(rng.choice([1 / 40, 1 / 12, 1 / 5])),
                            frac_ins=float(rng.choice([0, 0.2])), frac_del=float(rng.choice([0, 0.2])), n_runs=int(rng.choice([0, 2])),
                            lowercase_frac=float(rng.choice([0, 0.1])), two_beds=bool(rng.integers(0, 2)),
                            same_pos_frac=float(rng.choice([0, 0.05])))
    bgzf = bool(rng.integers(0, 2))
    a = fw.cohort_to_files(blk, pats, work, multiallelic_every=int(rng.choice([0, 7])), bgzf=bgzf, member_bytes=int(rng.choice([700, 4000, 60000])),
                           flank_records=int(rng.choice([0, 30])) if bgzf else 0, write_csi=bgzf and bool(rng.integers(0, 2)))
    kw, extra = {}, ["--chunk", str(int(rng.choice([1, 3, 50]))), "--threads", str(int(rng.choice([1, 3])))]
    if rng.random() < 0.3:
        kw["min_maf"] = int(rng.integers(1, 4))
        extra += ["--min_maf", str(kw["min_maf"])]
    if rng.random() < 0.3:
        kw["forward_only"] = True
        extra += ["--forward_only"]
    if rng.random() < 0.2:
        extra += ["--devices", "0,0"]
    try:
        expected = ora.run(a["chromosome"], a["bcf"], a["beds"], a["reference"], None, a["pwm_file"], a["threshold_dir"], 1e-4, a["names"], **kw)
    except ora.OracleError:
        return  # the input panics in the reference too
    out = os.path.join(work, "out.vcf.gz")
    cmd = [DRIVER, "--chromosome", a["chromosome"], "--input", a["bcf"], "--output", out, "--reference", a["reference"], "--bed", ",".join(a["beds"]),
           "--pwm_names", ",".join(a["names"]), "--pwm_file", a["pwm_file"], "--pwm_threshold_directory", a["threshold_dir"],
           "--pwm_threshold", "0.0001"] + extra
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-800:]
    assert ora.gunzip_file(out) == expected, "VCF differs"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--first-seed", type=int, default=0)
    args = ap.parse_args()
    t0, seed, bad = time.time(), args.first_seed, []
    while time.time() - t0 < args.seconds:
        work = tempfile.mkdtemp(prefix="tfbsfuzz")
        try:
            one_case(seed, work)
        except ValueError:
            pass  # generator limits (regions too small for a nested second BED set)
        except Exception as e:
            bad.append(seed)
            print("seed %d FAILED: %s" % (seed, str(e)[-600:]), flush=True)
        finally:
            shutil.rmtree(work, ignore_errors=True)
        seed += 1
    print("cases %d..%d: %d failures %s" % (args.first_seed, seed - 1, len(bad), bad))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
