"""The C++ driver (reference CLI on top of the C ABI) against the golden VCFs and against the oracle's run() on file sets."""
import os
import subprocess

import numpy as np
import pytest

from find_tfbs_b200 import synth
from oracle import pyoracle as ora
import file_writers as fw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.environ.get("TFBS_B200_DRIVER") or os.path.join(ROOT, "find_tfbs_b200", "find-tfbs-b200")  # the override is tests/cuda_emu


def run_driver(args, out):
    cmd = [DRIVER, "--chromosome", args["chromosome"], "--input", args["bcf"], "--output", out, "--reference", args["reference"],
           "--bed", ",".join(args["beds"]), "--pwm_names", ",".join(args["names"]), "--pwm_file", args["pwm_file"],
           "--pwm_threshold_directory", args["threshold_dir"], "--pwm_threshold", "0.0001"] + args.get("extra", [])
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return p


def golden_args(golden_dir, bcf):
    return {"chromosome": "chr1", "bcf": os.path.join(golden_dir, bcf), "reference": os.path.join(golden_dir, "reference_genome.fa"),
            "beds": [os.path.join(golden_dir, "regions1.bed"), os.path.join(golden_dir, "regions2.bed")], "names": ["ACGT"],
            "pwm_file": os.path.join(golden_dir, "pwm_definitions.txt"), "threshold_dir": golden_dir,
            "extra": ["--samples", os.path.join(golden_dir, "samples")]}


def test_file_writers_roundtrip_through_oracle_reader(tmp_path):
    """CPU: the BCF / FASTA writers used by the driver tests are read back exactly by the oracle's readers."""
    pats = synth.make_pwms(2, seed=3, lmin=6, lmax=10)
    blk = synth.make_cohort(5, 8, seed=3, lmax_pattern=10, region_len=(50, 120), variant_rate=0.05, two_beds=True)
    a = fw.cohort_to_files(blk, pats, str(tmp_path), multiallelic_every=7)
    b = ora.read_bcf(a["bcf"])
    assert b["samples"] == a["samples"] and b["contigs"] == ["chrOther", "chrS", "chrZ"]
    biallelic = [i for i, al in enumerate(b["alleles"]) if len(al) == 2]
    assert [b["pos"][i] for i in biallelic] == blk.meta["var_pos"].tolist()
    bits = np.unpackbits(blk.carriers.view(np.uint8), axis=1, bitorder="little")[:, :10]
    assert np.array_equal((b["gt"][biallelic, :, 0] == 4), bits[:len(biallelic), 0::2] == 1)
    assert np.array_equal((b["gt"][biallelic, :, 1] == 5), bits[:len(biallelic), 1::2] == 1)
    # the oracle's whole-program path on the files equals its block path on the arrays
    text = ora.run(a["chromosome"], a["bcf"], a["beds"], a["reference"], None, a["pwm_file"], a["threshold_dir"], 1e-4, a["names"])
    assert text.startswith("#CHROM\tPOS") and text.splitlines()[0].split("\t")[9:] == a["samples"]
    pw = ora.parse_pwm_files(a["pwm_file"], a["threshold_dir"], 1e-4, a["names"], True)
    assert len(pw) == len(pats)
    for x, y in zip(pw, pats):
        assert np.array_equal(x["weights"], y["weights"]) and x["min_score"] == y["min_score"] and x["pattern_id"] == y["pattern_id"]


@pytest.mark.gpu
def test_driver_golden_vcfs(golden_dir, tmp_path):
    """main.rs:548-568: both integration outputs, compared after gunzip."""
    for bcf, exp in (("genotypes.bcf", "expected_output_1.vcf.gz"), ("genotypes2.bcf", "expected_output_2.vcf.gz")):
        out = str(tmp_path / ("out_" + exp))
        p = run_driver(golden_args(golden_dir, bcf), out)
        assert p.returncode == 0, p.stderr
        assert ora.gunzip_file(out) == ora.gunzip_file(os.path.join(golden_dir, exp))
        assert not os.path.exists(out + ".part")


@pytest.mark.gpu
def test_driver_matches_oracle_on_synthetic_files(tmp_path):
    pats = synth.make_pwms(6, seed=13, lmin=6, lmax=22)
    lmax = max(p["weights"].shape[0] for p in pats)
    blk = synth.make_cohort(12, 40, seed=13, lmax_pattern=lmax, region_len=(80, 500), variant_rate=1 / 15.0, frac_ins=0.1, frac_del=0.1,
                            two_beds=True, n_runs=3, lowercase_frac=0.05, same_pos_frac=0.03)
    a = fw.cohort_to_files(blk, pats, str(tmp_path), multiallelic_every=11)
    expected = ora.run(a["chromosome"], a["bcf"], a["beds"], a["reference"], None, a["pwm_file"], a["threshold_dir"], 1e-4, a["names"])
    assert len(expected.splitlines()) > 5
    for extra in ([], ["--chunk", "7"], ["--plain"], ["--chunk", "5", "--threads", "3"], ["--chunk", "4", "--devices", "0,0"]):  # last: two contexts, one per worker thread
        out = str(tmp_path / "out.vcf.gz")
        a["extra"] = extra
        p = run_driver(a, out)
        assert p.returncode == 0, p.stderr
        got = open(out).read() if "--plain" in extra else ora.gunzip_file(out)
        assert got == expected
    # the same cohort inside a multi-contig BGZF file with a CSI index (read through the index) and with --no_index (whole-file scan)
    ix = fw.cohort_to_files(blk, pats, str(tmp_path / "indexed"), multiallelic_every=11, bgzf=True, member_bytes=3000, flank_records=25, write_csi=True)
    assert ora.run(ix["chromosome"], ix["bcf"], ix["beds"], ix["reference"], None, ix["pwm_file"], ix["threshold_dir"], 1e-4, ix["names"]) == expected
    for extra in ([], ["--no_index"]):
        out = str(tmp_path / "out_ix.vcf.gz")
        ix["extra"] = extra
        p = run_driver(ix, out)
        assert p.returncode == 0, p.stderr
        assert ora.gunzip_file(out) == expected
    # options: --min_maf, --forward_only, --after_position, --samples subset
    sub = str(tmp_path / "subset.txt")
    open(sub, "w").write("\n".join(a["samples"][1::2]) + "\n")
    for extra, kw in ((["--min_maf", "3"], {"min_maf": 3}), (["--forward_only"], {"forward_only": True}),
                      (["--after_position", str(blk.meta["merged"][10][0])], {"after_position": blk.meta["merged"][10][0]}),
                      (["--samples", sub], {})):
        out = str(tmp_path / "out2.vcf.gz")
        a["extra"] = extra
        p = run_driver(a, out)
        assert p.returncode == 0, p.stderr
        exp = ora.run(a["chromosome"], a["bcf"], a["beds"], a["reference"], sub if "--samples" in extra else None, a["pwm_file"],
                      a["threshold_dir"], 1e-4, a["names"], **kw)
        assert ora.gunzip_file(out) == exp


@pytest.mark.gpu
def test_driver_audit_file(golden_dir, tmp_path):
    """--audit: the VCF is unchanged, the fixture (ACGT scores 4000 > 3999, next best 3000) has no tie, no truncated haplotype;
    a synthetic file set with overlapping records lists its truncated haplotypes."""
    out, aud = str(tmp_path / "o.vcf.gz"), str(tmp_path / "audit.tsv")
    a = golden_args(golden_dir, "genotypes2.bcf")
    a["extra"] += ["--audit", aud]
    p = run_driver(a, out)
    assert p.returncode == 0, p.stderr
    assert ora.gunzip_file(out) == ora.gunzip_file(os.path.join(golden_dir, "expected_output_2.vcf.gz"))
    assert [ln for ln in open(aud).read().splitlines() if not ln.startswith("#")] == []
    pats = synth.make_pwms(3, seed=19, lmin=6, lmax=12)
    blk = synth.make_cohort(8, 20, seed=19, lmax_pattern=12, region_len=(80, 300), variant_rate=1 / 8.0, frac_del=0.3, same_pos_frac=0.1)
    f = fw.cohort_to_files(blk, pats, str(tmp_path))
    expected = ora.run(f["chromosome"], f["bcf"], f["beds"], f["reference"], None, f["pwm_file"], f["threshold_dir"], 1e-4, f["names"])
    f["extra"] = ["--audit", aud, "--chunk", "6"]
    p = run_driver(f, out)
    assert p.returncode == 0, p.stderr
    assert ora.gunzip_file(out) == expected
    lines = [ln.split("\t") for ln in open(aud).read().splitlines() if not ln.startswith("#")]
    assert any(ln[0] == "truncated" for ln in lines)
    assert all(ln[0] in ("tie", "truncated", "overwritten") for ln in lines)


@pytest.mark.gpu
def test_driver_reports_reference_panics(golden_dir, tmp_path):
    bad = golden_args(golden_dir, "genotypes2.bcf")
    bad["names"] = ["NOPE"]
    p = run_driver(bad, str(tmp_path / "x.vcf.gz"))
    assert p.returncode != 0 and "Could not open file" in p.stderr
