"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/tfbs.h declares, the
ctypes mirrors agree with each other, and the host-side mirrors of the reference's region logic agree with the oracle.
No compute entry point is called here (there is no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from find_tfbs_b200 import binding, synth
from oracle import pyoracle as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "tfbs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tfbs_[a-z_]+)\s*\(", text)))


def test_header_functions_are_exported():
    binding.build()
    L = C.CDLL(binding.LIB_PATH)
    names = declared_functions()
    assert set(names) == set(binding.EXPORTS)
    for n in names:
        assert hasattr(L, n), n
    L.tfbs_abi_version.restype = C.c_int
    assert L.tfbs_abi_version() == 2


def test_no_gpu_fails_loudly():
    """Without a usable device tfbs_create must fail with TFBS_ERR_CUDA: there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(binding.TfbsError) as e:
        binding.Context(0)
    assert e.value.code == binding.ERR_CUDA
    assert "no CPU fallback" in e.value.message


def test_struct_layouts_agree():
    for a, b in ((binding.TfbsPattern, ora.TfbsPattern), (binding.TfbsInnerRegion, ora.TfbsInnerRegion),
                 (binding.TfbsVariant, ora.TfbsVariant), (binding.TfbsBlock, ora.TfbsBlock)):
        assert C.sizeof(a) == C.sizeof(b)
        assert [(f[0], getattr(a, f[0]).offset) for f in a._fields_] == [(f[0], getattr(b, f[0]).offset) for f in b._fields_]
    assert binding.INNER_DTYPE.itemsize == C.sizeof(binding.TfbsInnerRegion) == 24
    assert binding.VARIANT_DTYPE.itemsize == C.sizeof(binding.TfbsVariant) == 32
    assert C.sizeof(binding.TfbsPattern) == 24


def test_product_does_not_touch_the_oracle():
    """Nothing under find_tfbs_b200/ may import, link or call oracle/."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "find_tfbs_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text and "oracle.hpp" not in text, os.path.join(dp, f)


def test_host_region_logic_matches_oracle(golden_dir):
    """merge_regions / select_inner_peaks mirrors (bed.rs:25-47, range.rs:43-87, main.rs:62-72) against the oracle."""
    beds = [os.path.join(golden_dir, "regions1.bed"), os.path.join(golden_dir, "regions2.bed")]
    merged, pm = ora.load_peak_files(beds, "chr1", 0)
    flat = [r for k in sorted(pm) for r in pm[k]]
    assert synth.merge_regions(flat) == merged
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(1, 30))
        s = rng.integers(0, 300, size=n)
        regs = [(int(a), int(a + rng.integers(0, 40))) for a in s]
        got = synth.merge_regions(regs)
        # oracle through temporary BED files
        import tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False) as f:
            for a, b in regs:
                f.write("chrT\t%d\t%d\n" % (a, b))
        m2, _ = ora.load_peak_files([f.name], "chrT", 0)
        os.unlink(f.name)
        assert got == m2
        for m in got:
            for p in regs:
                assert synth.range_overlaps(p, m) == ora.range_overlaps(p, m)
    # the strictly-inside region of the fixture is never selected
    sel = synth.select_inner_peaks((180, 210), [pm["regions1.bed"], pm["regions2.bed"]])
    assert [(x[0], x[1], x[2]) for x in sel] == [(180, 190, 0), (200, 210, 0)]


def test_synth_is_seeded_and_consistent():
    pats = synth.make_pwms(3, seed=7)
    pats2 = synth.make_pwms(3, seed=7)
    assert all(np.array_equal(a["weights"], b["weights"]) and a["min_score"] == b["min_score"] for a, b in zip(pats, pats2))
    assert np.array_equal(pats[1]["weights"], ora.reverse_complement(pats[0]["weights"]))
    a = synth.make_cohort(5, 10, seed=3, lmax_pattern=12, two_beds=True, n_runs=2)
    b = synth.make_cohort(5, 10, seed=3, lmax_pattern=12, two_beds=True, n_runs=2)
    assert np.array_equal(a.ref_bases, b.ref_bases) and np.array_equal(a.carriers, b.carriers) and np.array_equal(a.variants, b.variants)
    # REF alleles are taken from the genome: first base of every in-window record matches the window
    for r in range(a.n_regions):
        for v in a.variants[a.var_off[r]:a.var_off[r + 1]]:
            if a.region_start[r] <= v["pos"] <= a.region_end[r]:
                i = int(a.ref_off[r]) + int(v["pos"] - a.region_start[r])
                assert chr(a.ref_bases[i]).upper() == chr(a.allele_bases[v["ref_off"]]).upper()


def test_score_threshold_tail():
    """score_threshold returns the largest s with P(score >= s) > p; brute force on a short PWM."""
    rng = np.random.default_rng(1)
    w = rng.integers(-500, 500, size=(5, 4))
    import itertools
    scores = np.array([sum(w[c, b] for c, b in enumerate(t)) for t in itertools.product(range(4), repeat=5)])
    for p in (0.3, 0.05, 0.004):
        s = synth.score_threshold(w, p)
        assert (scores >= s).mean() > p and (scores >= s + 1).mean() <= p


def test_oracle_small_synthetic_runs_and_is_thread_invariant():
    import parity_helpers as hp
    pats = synth.make_pwms(3, seed=5, lmin=5, lmax=12)
    blk = synth.make_cohort(6, 12, seed=5, lmax_pattern=12, region_len=(60, 200), two_beds=True, same_pos_frac=0.1, variant_rate=0.1)
    ps = binding.PatternSet(pats)
    a = hp.run_oracle(ps, blk, 1, True, 1)
    b = hp.run_oracle(ps, blk, 1, True, 5)
    for k in ("region", "inner", "pattern_id", "left", "right", "m_start", "hap_group"):
        assert np.array_equal(a[k], b[k])
    assert a["executed_cells"] == b["executed_cells"] and a["n_hits"] > 0
