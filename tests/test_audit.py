"""CPU checks of the audit's definition: the difference of two hit lists (thresholds lowered by one) is exactly the set of windows
scoring min_score; checked for the reference haplotype by brute force with numpy.  The CUDA side of the audit (tfbs_audit_block)
is compared with these sets in tests/test_gpu_parity.py."""
import numpy as np

from find_tfbs_b200 import synth
from find_tfbs_b200.binding import PatternSet
import parity_helpers as hp

CODE = np.full(256, 4, dtype=np.int64)
for i, ch in enumerate("ACGT"):
    CODE[ord(ch)] = CODE[ord(ch.lower())] = i


def test_ties_are_windows_scoring_exactly_min_score():
    rng = np.random.default_rng(3)
    w = rng.integers(-3, 4, size=(8, 4)).astype(np.int32) * 100
    best = int(w.max(axis=1).sum())
    pats = [{"weights": w, "min_score": best - 400, "pattern_id": 1, "direction": 0},
            {"weights": w[::-1, ::-1].copy(), "min_score": best - 300, "pattern_id": 1, "direction": 1}]
    ps = PatternSet(pats)
    blk = synth.make_cohort(6, 16, seed=9, lmax_pattern=8, region_len=(150, 400), n_runs=2)
    au = hp.oracle_audit(ps, blk)
    ref_ties = set(t for t in au["ties"] if t[2] == -1)
    assert ref_ties
    # brute force over the reference windows of every region whose reference group has members (main.rs:129)
    b = hp.run_oracle(ps, blk, 1, True)
    hg = b["hap_group"].reshape(blk.n_regions, -1)
    want = set()
    for r in range(blk.n_regions):
        if not (hg[r] == 0).any():
            continue
        codes = CODE[blk.ref_bases[int(blk.ref_off[r]):int(blk.ref_off[r + 1])]]
        for pi, p in enumerate(pats):
            wt = np.concatenate([p["weights"].astype(np.int64), np.zeros((len(p["weights"]), 1), dtype=np.int64)], axis=1)  # N scores 0
            L = wt.shape[0]
            for i in range(len(codes) - L + 1):
                if int(wt[np.arange(L), codes[i:i + L]].sum()) == p["min_score"]:
                    want.add((r, pi, -1, int(blk.region_start[r]) + i))
    assert ref_ties == want


def test_flags_mark_truncated_and_overwritten_haplotypes():
    """Two overlapping deletions carried together truncate (haplotype.rs:144-149); a haplotype that carries a record outside the window
    next to an in-window one patches to the same sequence as the in-window record alone: one of the two entries is overwritten
    (haplotype.rs:84) and its haplotypes are counted with the reference."""
    ref = "ACGTACGTACGTACGTACGT"
    blk = hp.hand_block(3, [(10, 29, ref)],
                        [(0, 12, "GTA", "G"), (0, 13, "T", "C"), (0, 5, "A", "C"), (0, 20, "G", "T")],
                        [[0], [0, 1], [3], [2, 3]])
    ps = PatternSet([{"weights": np.eye(4, dtype=np.int32)[[0, 1]] * 10, "min_score": 15, "pattern_id": 0}])
    o = hp.run_oracle(ps, blk, 1, True)
    fl = o["hap_flags"]
    assert fl[0] & 1            # del GTA->G at 12 then SNV at 13 (< cursor 15): truncated
    assert not fl[1] & 1        # the SNV alone is fine
    assert fl[3] == 2 and fl[2] == 0   # {out-of-window 5, SNV 20} patches like {SNV 20}: the later group is overwritten
    assert o["hap_group"][3] == 0 and o["hap_group"][2] != 0
    assert o["collision_regions"] == 1
