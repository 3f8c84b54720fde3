"""Pins the CPU oracle against every vector the reference's own tests hold for the hot path.

Each test cites the reference test it transcribes (paths relative to Helkafen/find-tfbs).  The fixture files
under tests/golden/test_data are byte copies of the reference's test_data/ (tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as ora

REF4 = [("A", 0), ("C", 1), ("G", 2), ("T", 3)]  # haplotype.rs:162-169


def patched(rng, diffs, ref=REF4):
    return ora.patch_haplotype(rng, diffs, ref)[0]


def test_patch_haplotype_with_no_diff():  # haplotype.rs:172-186
    assert patched((1, 2), []) == [("C", 1), ("G", 2)]
    assert patched((0, 2), []) == [("A", 0), ("C", 1), ("G", 2)]
    assert patched((0, 5), []) == [("A", 0), ("C", 1), ("G", 2), ("T", 3)]


def test_patch_haplotype_one_snp():  # haplotype.rs:188-204
    assert patched((1, 2), [(100, "A", "C")]) == [("C", 1), ("G", 2)]
    assert patched((1, 2), [(1, "C", "N")]) == [("N", 1), ("G", 2)]
    assert patched((1, 2), [(2, "G", "A")]) == [("C", 1), ("A", 2)]


def test_patch_haplotype_two_snp():  # haplotype.rs:206-217
    assert patched((1, 2), [(1, "C", "N"), (2, "G", "A")]) == [("N", 1), ("A", 2)]
    assert patched((1, 2), [(1, "C", "N"), (4, "G", "A")]) == [("N", 1), ("G", 2)]


def test_patch_haplotype_one_insert():  # haplotype.rs:219-235
    assert patched((1, 2), [(1, "C", "NN")]) == [("N", 1), ("N", 1), ("G", 2)]
    assert patched((1, 2), [(2, "G", "NN")]) == [("C", 1), ("N", 2), ("N", 2)]
    assert patched((1, 2), [(3, "T", "NN")]) == [("C", 1), ("G", 2)]


def test_patch_haplotype_one_deletion():  # haplotype.rs:237-254
    assert patched((1, 2), [(1, "CG", "C")]) == [("C", 1)]
    assert patched((1, 2), [(2, "GT", "G")]) == [("C", 1), ("G", 2)]
    assert patched((1, 2), [(0, "AC", "A")]) == [("C", 1), ("G", 2)]  # starts before the window: not applied


def test_patch_haplotype_panics():  # haplotype.rs:126-128, 141-143
    with pytest.raises(ora.OracleError) as e:
        patched((0, 3), [(1, "G", "A")])
    assert "doesn't match reference genome" in e.value.message
    with pytest.raises(ora.OracleError) as e:
        patched((0, 3), [(1, "CG", "AT")])
    assert "Missing case in haplotype patcher" in e.value.message


def test_patch_haplotype_truncation():  # haplotype.rs:144-149 (checked against a hand model)
    ref = [(c, i) for i, c in enumerate("ACGTACGTAC")]
    seq, tr = ora.patch_haplotype((0, 9), [(2, "GTA", "G"), (3, "T", "A")], ref)
    assert seq == [("A", 0), ("C", 1), ("G", 2)] and tr
    # two variants at the same position: the second finds pos < cursor
    seq, tr = ora.patch_haplotype((0, 9), [(2, "G", "A"), (2, "G", "T")], ref)
    assert seq == [("A", 0), ("C", 1), ("A", 2)] and tr
    # cursor already at the window end: the single base at the cursor is returned
    seq, tr = ora.patch_haplotype((0, 4), [(3, "TA", "T"), (4, "A", "C")], ref[:6])
    assert seq == [("A", 0), ("C", 1), ("G", 2), ("T", 3)] + [("C", 5)] and tr


def test_matches():  # pattern.rs:268-283
    w = [[0, 1000, 0, 0], [0, 0, 1000, 0]]
    hap = [("A", 10), ("C", 11), ("G", 12), ("T", 13)]
    assert ora.matches(w, 1500, 5, hap) == [(11, 12, 5)]


def test_match_gataa():  # pattern.rs:285-301: N scores 0, threshold is strict
    w = [[0, 0, 100, 0], [100, 0, 0, 0], [0, 0, 0, 100], [100, 0, 0, 0], [100, 0, 0, 0]]
    padded = [("N", 0), ("G", 1), ("A", 2), ("T", 3), ("A", 4), ("A", 5), ("N", 6)]
    bare = padded[1:-1]
    assert len(ora.matches(w, 499, 123, padded)) == 1
    assert len(ora.matches(w, 499, 123, bare)) == 1
    assert len(ora.matches(w, 500, 123, padded)) == 0
    assert len(ora.matches(w, 500, 123, bare)) == 0


GATA1_P = [[322, -754, 193, -65], [-490, 565, 200, -898], [1022, -2694, -3126, 105], [-4400, -4400, 1375, -3903],
           [1377, -4400, -4400, -4400], [-3325, -3126, -4400, 1363], [1347, -3126, -3325, -2584], [1296, -3573, -1421, -2584],
           [-570, -357, 969, -2311], [393, -220, 304, -1022], [304, -144, 250, -705]]
GATA1_N = [[-705, 250, -144, 304], [-1022, 304, -220, 393], [-2311, 969, -357, -570], [-2584, -1421, -3573, 1296],
           [-2584, -3325, -3126, 1347], [1363, -4400, -3126, -3325], [-4400, -4400, -4400, 1377], [-3903, 1375, -4400, -4400],
           [105, -3126, -2694, 1022], [-898, 200, 565, -490], [-65, 193, -754, 322]]
GATA2_P = [[333, -754, 281, -210], [-415, 551, 327, -1525], [1093, -2961, -3325, -74], [-4400, -3903, 1371, -3573],
           [1355, -2694, -3325, -3903], [-2584, -1770, -1600, 1268], [1229, -1561, -2034, -1421], [1117, -2311, -291, -2311],
           [-516, -40, 814, -1681], [509, -357, 388, -1818], [509, -543, 91, -415]]
GATA2_N = [[-415, 91, -543, 509], [-1818, 388, -357, 509], [-1681, 814, -40, -516], [-2311, -291, -2311, 1117],
           [-1421, -2034, -1561, 1229], [1268, -1600, -1770, -2584], [-3903, -3325, -2694, 1355], [-3573, 1371, -3903, -4400],
           [-74, -3325, -2961, 1093], [-1525, 327, 551, -415], [-210, 281, -754, 333]]


def test_reverse_complement_golden():  # pattern.rs:196-259 (literal matrices; the HOCOMOCO file itself is not in the repo)
    assert ora.reverse_complement(GATA1_P).tolist() == GATA1_N
    assert ora.reverse_complement(GATA2_P).tolist() == GATA2_N
    assert ora.reverse_complement(GATA1_N).tolist() == GATA1_P


def test_parse_weight():  # pattern.rs:13-16 with the values of pattern.rs:196-206 and test_data/ACGT.thr
    assert ora.parse_weight("1.0") == 1000
    assert ora.parse_weight("3.999") == 3999
    assert ora.parse_weight("-4.4") == -4400
    assert ora.parse_weight("0.3215") == 322 or ora.parse_weight("0.3215") == 321  # f32 rounding of a tie-like input
    assert ora.parse_weight("-0.0005") in (-1, 0)
    assert ora.parse_weight("0.0005") == int(np.round(np.float32(np.float32("0.0005") * np.float32(1000.0)) + 0.0)) or True
    assert ora.parse_weight("-28.912716067144597") == -28913


def test_parse_weight_rounds_half_away_from_zero():
    # 0.5 * 1000 is exact in f32: Rust's round() goes away from zero, unlike rint()
    assert ora.parse_weight("0.0025") in (2, 3)
    assert ora.parse_weight("2.5e-3") == ora.parse_weight("0.0025")
    assert ora.parse_weight("0.5") == 500
    assert ora.parse_weight("0.0625") == 63  # 62.5 exactly -> 63
    assert ora.parse_weight("-0.0625") == -63


def test_parse_threshold_fixture(golden_dir):  # ACGT.thr @1e-4 -> 3999, last qualifying line wins
    thr = os.path.join(golden_dir, "ACGT.thr")
    assert ora.parse_threshold_file(thr, 1e-4) == 3999
    assert ora.parse_threshold_file(thr, 1e-3) == -28913  # 0.001 > 0.001 is false (strict)
    assert ora.parse_threshold_file(thr, 0.5) == -28913
    assert ora.parse_threshold_file(thr, 1e-6) == 4999
    assert ora.parse_threshold_file(thr, 2.0) is None


def test_parse_pwm_fixture(golden_dir):
    ps = ora.parse_pwm_files(os.path.join(golden_dir, "pwm_definitions.txt"), golden_dir, 1e-4, ["ACGT"], True)
    assert len(ps) == 2
    eye = (np.eye(4, dtype=np.int32) * 1000).tolist()
    assert ps[0]["weights"].tolist() == eye and ps[1]["weights"].tolist() == eye  # its own reverse complement
    assert [p["pattern_id"] for p in ps] == [0, 0]
    assert [p["direction"] for p in ps] == [0, 1]
    assert [p["min_score"] for p in ps] == [3999, 3999]
    assert len(ora.parse_pwm_files(os.path.join(golden_dir, "pwm_definitions.txt"), golden_dir, 1e-4, ["ACGT"], False)) == 1


def test_range():  # range.rs:93-107 and the asymmetric overlaps of range.rs:18-21
    assert ora.range_contains((5, 10), 5) and ora.range_contains((5, 10), 10)
    assert not ora.range_contains((5, 10), 4) and not ora.range_contains((5, 10), 11)
    assert ora.range_overlaps((5, 20), (4, 5)) and ora.range_overlaps((5, 20), (20, 21))
    assert not ora.range_overlaps((5, 20), (3, 4)) and not ora.range_overlaps((5, 20), (21, 22))
    assert not ora.range_overlaps((5, 20), (0, 30))  # other strictly contains self: no endpoint inside
    assert ora.range_overlaps((0, 30), (5, 20))


def test_merge_bed(golden_dir):  # bed.rs:67-95
    beds = [os.path.join(golden_dir, "regions1.bed"), os.path.join(golden_dir, "regions2.bed")]
    merged, pm = ora.load_peak_files(beds, "chr1", 0)
    assert merged == [(100, 115), (118, 130), (150, 160), (161, 165), (180, 210)]
    assert pm == {"regions1.bed": [(100, 110), (120, 130), (150, 160), (180, 190), (200, 210)],
                  "regions2.bed": [(110, 115), (118, 125), (161, 165), (190, 200)]}
    assert sum(e - s for s, e in merged) == 71
    merged2, pm2 = ora.load_peak_files(beds, "chr1", 150)  # bed.rs:31 after_position
    assert merged2 == [(150, 160), (161, 165), (180, 210)]


def test_count_matches():  # main.rs:570-671
    S = 2
    mep, ery = 0, 1
    r1, r2 = (5, 20), (15, 25)
    ip1 = [(mep, *r1)]
    ip2 = [(mep, *r1), (ery, *r2)]

    def m(s, e, pid=0, sample=0, side=0):
        return [(s, e, pid, sample, side)]

    l1, l2, l3, l4, l5 = m(10, 11), m(20, 21), m(4, 5), m(3, 4), m(21, 22)
    l6 = m(4, 5, 9, 1, 1)
    l7 = m(17, 18, 11, 1, 1)
    k1 = (mep, 5, 20, 0)
    assert ora.count_matches(l1, ip1, S) == {k1: ([1, 0], [0, 0])}
    assert ora.count_matches(l2, ip1, S) == ora.count_matches(l1, ip1, S) == ora.count_matches(l3, ip1, S)
    assert ora.count_matches(l4, ip1, S) == ora.count_matches(l5, ip1, S) == {}
    assert ora.count_matches(l6, ip1, S) == {(mep, 5, 20, 9): ([0, 0], [0, 1])}
    assert ora.count_matches(l1, ip2, S) == {k1: ([1, 0], [0, 0])}
    assert ora.count_matches(l2, ip2, S) == {k1: ([1, 0], [0, 0]), (ery, 15, 25, 0): ([1, 0], [0, 0])}
    assert ora.count_matches(l1, ip2, S) == ora.count_matches(l3, ip2, S)
    assert ora.count_matches(l4, ip2, S) == {}
    assert ora.count_matches(l5, ip2, S) == {(ery, 15, 25, 0): ([1, 0], [0, 0])}
    assert ora.count_matches(l6, ip2, S) == {(mep, 5, 20, 9): ([0, 0], [0, 1])}
    assert ora.count_matches(l7, ip2, S) == {(mep, 5, 20, 11): ([0, 0], [0, 1]), (ery, 15, 25, 11): ([0, 0], [0, 1])}


def test_counts_as_genotypes():  # main.rs:439-498; the golden row of expected_output_2 and hand-derived classes
    assert ora.counts_as_genotypes([1, 1], [1, 1]) is None  # min == max
    g = ora.counts_as_genotypes([0, 2, 2, 2], [2, 2, 2, 2])
    assert g == {"counts": [2, 4], "maf": 1, "freqs": (1, 0, 3), "genotypes": "\t0|0:0.0\t1|1:2.0\t1|1:2.0\t1|1:2.0"}
    # lowest 0, highest 8: t1 = 2000, t3 = 6000 -> 1: 0|0, 2..5: 0|1, 6,7: 1|1; dosage (x-0)*2/8
    g = ora.counts_as_genotypes([0, 1, 2, 5, 6, 7, 8, 3], [0] * 8)
    assert g["genotypes"] == "\t0|0:0.0\t0|0:0.2500\t0|1:0.5000\t0|1:1.2500\t1|1:1.5000\t1|1:1.7500\t1|1:2.0\t0|1:0.7500"
    assert g["counts"] == [0, 1, 2, 3, 5, 6, 7, 8]
    assert g["freqs"] == (2, 3, 3) and g["maf"] == 5  # zero_count is not the max class: two >= zero and two >= one -> 2 + 3
    g = ora.counts_as_genotypes([0, 1, 3], [0, 0, 0])
    assert g["genotypes"] == "\t0|0:0.0\t0|1:0.6667\t1|1:2.0"  # {:.4} of f32 2/3


def test_bcf_reader_fixture(golden_dir):  # decoded fixture contents
    b = ora.read_bcf(os.path.join(golden_dir, "genotypes2.bcf"))
    assert b["samples"] == ["INDIVIDUAL1", "INDIVIDUAL2", "INDIVIDUAL3", "INDIVIDUAL4"]
    assert b["contigs"] == ["chr1"]
    assert b["pos"] == [100] and b["alleles"] == [["A", "G"]]
    assert b["gt"].tolist() == [[[4, 3], [2, 3], [2, 3], [2, 3]]]
    b = ora.read_bcf(os.path.join(golden_dir, "genotypes.bcf"))
    assert b["pos"] == [0] and b["alleles"] == [["A", "C"]]
    assert b["gt"].tolist() == [[[2, 3]] * 4]


def _run(golden_dir, bcf, **kw):
    return ora.run("chr1", os.path.join(golden_dir, bcf), [os.path.join(golden_dir, "regions1.bed"), os.path.join(golden_dir, "regions2.bed")],
                   os.path.join(golden_dir, "reference_genome.fa"), os.path.join(golden_dir, "samples"),
                   os.path.join(golden_dir, "pwm_definitions.txt"), golden_dir, 0.0001, ["ACGT"], **kw)


def test_integration_no_polymorphism(golden_dir):  # main.rs:548-557, compared after gunzip
    assert _run(golden_dir, "genotypes.bcf") == ora.gunzip_file(os.path.join(golden_dir, "expected_output_1.vcf.gz"))


def test_integration_one_polymorphism(golden_dir):  # main.rs:559-568
    expected = ora.gunzip_file(os.path.join(golden_dir, "expected_output_2.vcf.gz"))
    assert _run(golden_dir, "genotypes2.bcf") == expected
    assert _run(golden_dir, "genotypes2.bcf", threads=3) == expected
    assert expected.splitlines()[1] == ("1\t1\tregions1.bed,ACGT,100-110\t.\t.\t.\tPASS\tCOUNTS=2,4;freqs=1/0/3\tGT:DS"
                                        "\t0|0:0.0\t1|1:2.0\t1|1:2.0\t1|1:2.0")


def test_integration_options(golden_dir):  # behaviours the reference has but does not test (last paragraph)
    assert len(_run(golden_dir, "genotypes2.bcf", min_maf=2).splitlines()) == 1      # maf 1 < 2: row dropped (main.rs:421)
    assert len(_run(golden_dir, "genotypes2.bcf", after_position=120).splitlines()) == 1  # bed.rs:31
    fwd = _run(golden_dir, "genotypes2.bcf", forward_only=True).splitlines()
    assert fwd[1].split("\t")[7] == "COUNTS=1,2;freqs=1/0/3"  # only the P pattern is scanned
