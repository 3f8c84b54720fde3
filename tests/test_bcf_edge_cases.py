"""BCF edge cases through the C++ driver, with expectations derived BY HAND from the reference's load_diffs (src/haplotype.rs:13-62)
and rust-htslib 0.26.1's Genotypes::get / GenotypeAllele::from -- not from the oracle's BCF reader (oracle/oracle_io.cpp shares an
author with the product's reader, so it cannot pin it).

The file is written byte by byte from the BCF2.2 specification: typed integer vectors of width 8 / 16 / 32 bits for FORMAT/GT,
END_OF_VECTOR padding of a haploid call, missing alleles, a first allele carrying the phase bit, a deletion record that starts in
front of the window, a symbolic ALT and a record without FORMAT/GT.

Derivation (reference semantics):
  * genotypes come back as i32 whatever the stored width (bcf_get_format_int32), allele code = (allele + 1) << 1 | phased;
  * left carries ALT iff value[0] is Unphased(1) = raw 4; right iff value[1] is Phased(1) = raw 5 (haplotype.rs:34-41); so `1|1`
    written 5,5 gives only the right haplotype, `0/1` = 2,4 gives nothing, `./.` = 0,0 and `.|.` = 0,1 give nothing;
  * a record is fetched when [pos, pos + rlen) overlaps [window.start, window.end + 1); a deletion that STARTS before the window is in
    the haplotype's diff list but patch_haplotype drops it (haplotype.rs:95, 249-253): the haplotype scores like the reference;
  * rust-htslib trims END_OF_VECTOR, so a haploid call has genotype.len() == 1 and `assert!(number_of_alleles == genotype.len())`
    (haplotype.rs:32) panics -- but only for records a region fetches; to_nucleotides on a symbolic ALT panics (util.rs:15), a record
    without GT makes record.genotypes().unwrap() panic (haplotype.rs:24) -- same condition.

Fixture: chr1 = 600 x 'A' with ACGT at 110, 120, 130, 140; one BED region 100-160; PWM ACGT (identity x 1000, threshold 3999, its own
reverse complement: a motif counts twice, pattern.rs:73-77); window = [97, 163].  Four samples:
  pos  95 AAAA->A (deletion, starts before the window)  S1 left
  pos 110 A->G  int16 GT   S0 4,3 (left)   S1 2,5 (right)   S2 5,5 (right only)   S3 2,4 (nobody)
  pos 120 A->G  int32 GT   S0 0,0 (none)   S1 0,1 (none)    S2 4,2 (left)         S3 4,5 (both)
  pos 130 A->G  int8 GT    S0 2,3          S1 2,3           S2 2,3                S3 2,5 (right)
intact motifs x 2:  S0 6 + 8 = 14, S1 8 + 6 = 14, S2 6 + 6 = 12, S3 6 + 4 = 10  ->  min 10, max 14, thresholds 11000 / 13000:
  1|1:2.0  1|1:2.0  0|1:1.0000  0|0:0.0   COUNTS=10,12,14  freqs=1/1/2   (main.rs:459-498)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import file_writers as fw
from oracle import pyoracle as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.environ.get("TFBS_B200_DRIVER") or os.path.join(ROOT, "find_tfbs_b200", "find-tfbs-b200")

SAMPLES = ["S0", "S1", "S2", "S3"]
EXPECTED_ROW = "1\t1\tregions.bed,ACGT,100-160\t.\t.\t.\tPASS\tCOUNTS=10,12,14;freqs=1/1/2\tGT:DS\t1|1:2.0\t1|1:2.0\t0|1:1.0000\t0|0:0.0"
END8, END16, END32 = -127, -32767, -2147483647  # END_OF_VECTOR of the three integer widths (BCF2.2 section 6.3.3)


def gt_block(width, pairs):
    """FORMAT/GT of one record: key = dictionary index 1 as a typed int8, then the vector descriptor (2 values of `width` bits per
    sample) and the values, little-endian."""
    code = {8: (0x21, "<b"), 16: (0x22, "<h"), 32: (0x23, "<i")}[width]
    out = bytes([0x11, 1, code[0]])
    for a, b in pairs:
        out += struct.pack(code[1], a) + struct.pack(code[1], b)
    return out


def record(pos, ref, alts, indiv, n_fmt=1):
    rlen = len(ref)
    shared = struct.pack("<iiiIII", 0, pos, rlen, 0x7F800001, ((1 + len(alts)) << 16), (n_fmt << 24) | len(SAMPLES))
    shared += bytes([0x07])  # ID: empty string
    for a in [ref] + alts:
        shared += fw._typed_str(a)
    shared += bytes([0x00])  # FILTER: empty vector
    return struct.pack("<II", len(shared), len(indiv)) + shared + indiv


def write_fixture(d, extra_records=()):
    genome = bytearray(b"A" * 600)
    for p in (110, 120, 130, 140):
        genome[p:p + 4] = b"ACGT"
    fa = os.path.join(d, "genome.fa")
    fw.write_fasta(fa, "chr1", bytes(genome), width=50)
    bed = os.path.join(d, "regions.bed")
    open(bed, "w").write("chr1\t100\t160\t1.0\n")
    pwm = os.path.join(d, "pwm.txt")
    open(pwm, "w").write(">ACGT\n1.0\t0.0\t0.0\t0.0\n0.0\t1.0\t0.0\t0.0\n0.0\t0.0\t1.0\t0.0\n0.0\t0.0\t0.0\t1.0\n")
    open(os.path.join(d, "ACGT.thr"), "w").write("-28.9\t1.0\n2.999\t0.001\n3.999\t0.00011\n4.999\t0.00001\n")
    text = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##contig=<ID=chr1,length=600>\n"
            "##ALT=<ID=DEL,Description=\"Deletion\">\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n" % "\t".join(SAMPLES)).encode() + b"\0"
    recs = [
        (95, record(95, "AAAA", ["A"], gt_block(8, [(2, 3), (4, 3), (2, 3), (2, 3)]))),
        (110, record(110, "A", ["G"], gt_block(16, [(4, 3), (2, 5), (5, 5), (2, 4)]))),
        (120, record(120, "A", ["G"], gt_block(32, [(0, 0), (0, 1), (4, 2), (4, 5)]))),
        (130, record(130, "A", ["G"], gt_block(8, [(2, 3), (2, 3), (2, 3), (2, 5)]))),
    ] + list(extra_records)
    recs.sort(key=lambda x: x[0])
    body = b"BCF\2\2" + struct.pack("<I", len(text)) + text + b"".join(r for _, r in recs)
    bcf = os.path.join(d, "edge.bcf")
    with open(bcf, "wb") as f:
        for i in range(0, len(body), 300):  # several BGZF members, records straddle them
            f.write(fw.bgzf_block(body[i:i + 300]))
        f.write(fw.bgzf_block(b""))
    return {"chromosome": "chr1", "bcf": bcf, "reference": fa, "bed": bed, "pwm": pwm, "thr": d}


def run(a, out):
    cmd = [DRIVER, "--chromosome", a["chromosome"], "--input", a["bcf"], "--output", out, "--reference", a["reference"], "--bed", a["bed"],
           "--pwm_names", "ACGT", "--pwm_file", a["pwm"], "--pwm_threshold_directory", a["thr"], "--pwm_threshold", "0.0001"]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


HAPLOID = lambda pos: (pos, record(pos, "A", ["C"], gt_block(8, [(4, END8), (2, 3), (2, 3), (2, 3)])))          # noqa: E731
HAPLOID16 = lambda pos: (pos, record(pos, "A", ["C"], gt_block(16, [(2, 3), (2, 3), (4, END16), (2, 3)])))      # noqa: E731
SYMBOLIC = lambda pos: (pos, record(pos, "A", ["<DEL>"], gt_block(8, [(2, 3), (2, 3), (2, 3), (2, 3)])))        # noqa: E731
NO_GT = lambda pos: (pos, record(pos, "A", ["C"], b"", n_fmt=0))                                                  # noqa: E731
NO_ALT = lambda pos: (pos, struct.pack("<II", 24 + 1 + 2 + 1, 0) + struct.pack("<iiiIII", 0, pos, 1, 0x7F800001, (1 << 16), len(SAMPLES)) +
                      bytes([0x07]) + fw._typed_str("A") + bytes([0x00]))                                         # noqa: E731


@pytest.mark.gpu
def test_gt_widths_phase_bits_missing_alleles_and_the_deletion_before_the_window(tmp_path):
    d = str(tmp_path)
    a = write_fixture(d)
    out = os.path.join(d, "out.vcf.gz")
    p = run(a, out)
    assert p.returncode == 0, p.stderr
    lines = ora.gunzip_file(out).strip().split("\n")
    assert lines[0].split("\t")[9:] == SAMPLES
    assert lines[1:] == [EXPECTED_ROW]


@pytest.mark.gpu
def test_records_the_reference_would_panic_on_only_matter_inside_a_region(tmp_path):
    """A haploid call, a symbolic ALT, a record without GT or without ALT: harmless anywhere outside the extended regions (the
    reference never fetches them), fatal inside one (haplotype.rs:21-32), with the reference's message."""
    for k, (make, msg) in enumerate(((HAPLOID, "Inconsistent number of alleles"), (HAPLOID16, "Inconsistent number of alleles"),
                                     (SYMBOLIC, "Unknown nucleotide 60"), (NO_GT, "missing GT"), (NO_ALT, "index out of bounds"))):
        d = str(tmp_path / ("far%d" % k))
        os.makedirs(d)
        out = os.path.join(d, "out.vcf.gz")
        p = run(write_fixture(d, [make(300), make(20)]), out)   # in front of and behind the only region's window [97, 163]
        assert p.returncode == 0, p.stderr
        assert ora.gunzip_file(out).strip().split("\n")[1:] == [EXPECTED_ROW]
        d = str(tmp_path / ("in%d" % k))
        os.makedirs(d)
        p = run(write_fixture(d, [make(150)]), os.path.join(d, "out.vcf.gz"))
        assert p.returncode == 101 and msg in p.stderr, (p.returncode, p.stderr)
        # the window is [97, 163]: 163 is inside, 164 is not
        d = str(tmp_path / ("edge%d" % k))
        os.makedirs(d)
        assert run(write_fixture(d, [make(164)]), os.path.join(d, "o.vcf.gz")).returncode == 0
        assert run(write_fixture(d, [make(163)]), os.path.join(d, "o.vcf.gz")).returncode == 101
