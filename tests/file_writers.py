"""Writers for the reference's input formats (FASTA + .fai, BED, HOCOMOCO-style PWM + .thr, BCF2.2 in gzip members),
used to push synthetic cohorts through the file-level entry points (the C++ driver and the oracle's run())."""
import gzip
import os
import struct

import numpy as np


def write_fasta(path, chrom, genome_bytes, width=60):
    seq = bytes(genome_bytes)
    with open(path, "wb") as f:
        hdr = (">%s\n" % chrom).encode()
        f.write(hdr)
        for i in range(0, len(seq), width):
            f.write(seq[i:i + width] + b"\n")
    with open(path + ".fai", "w") as f:
        f.write("%s\t%d\t%d\t%d\t%d\n" % (chrom, len(seq), len(hdr), width, width + 1))


def write_bed(path, chrom, regions, other_chrom_rows=2):
    with open(path, "w") as f:
        for s, e in regions:
            f.write("%s\t%d\t%d\t0.5\n" % (chrom, s, e))
        for k in range(other_chrom_rows):
            f.write("chrOther\t%d\t%d\t0.1\n" % (100 + k, 200 + k))


def write_pwms(pwm_path, thr_dir, pats):
    """pats: list from synth.make_pwms (forward entries carry the name); weights are written as w/1000 with 3 decimals so that
    parse_weight (f32 * 1000, round) gives the integers back."""
    os.makedirs(thr_dir, exist_ok=True)
    with open(pwm_path, "w") as f:
        for p in pats:
            if p["direction"] != 0:
                continue
            f.write(">%s\n" % p["name"])
            for row in p["weights"]:
                f.write("\t".join("%.3f" % (x / 1000.0) for x in row) + "\n")
            with open(os.path.join(thr_dir, p["name"] + ".thr"), "w") as t:
                t.write("%.3f\t1.0\n" % ((p["min_score"] - 5000) / 1000.0))
                t.write("%.3f\t0.00011\n" % (p["min_score"] / 1000.0))
                t.write("%.3f\t0.00001\n" % ((p["min_score"] + 2000) / 1000.0))


def _typed_str(s):
    b = s.encode()
    if len(b) < 15:
        return bytes([(len(b) << 4) | 7]) + b
    if len(b) < 128:
        return bytes([0xF7, 0x11, len(b)]) + b
    return bytes([0xF7, 0x12]) + struct.pack("<h", len(b)) + b


def bgzf_block(data):
    """One BGZF member: gzip header with the BC extra subfield (BSIZE), raw deflate, CRC32, ISIZE."""
    import zlib
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bcf(path, chrom, contig_len, samples, records, member_bytes=40000, bgzf=False, flank_records=0, write_csi=False):
    """records: list of (pos, [alleles...], gt) with gt an (n_samples, 2) int array of raw BCF codes, sorted by pos.
    flank_records: that many records of contig chrOther before and of contig chrZ after the wanted contig (a multi-contig file);
    write_csi: also write <path>.csi (one bin per contig holding one chunk = the contig's records), needs bgzf."""
    text = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##contig=<ID=chrOther,length=1000>\n"
            "##contig=<ID=%s,length=%d>\n##contig=<ID=chrZ,length=1000>\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n" % (chrom, contig_len, "\t".join(samples))).encode() + b"\0"
    out = bytearray(b"BCF\2\2" + struct.pack("<I", len(text)) + text)
    n = len(samples)
    flank_gt = np.tile(np.array([[4, 5]], dtype=np.int8), (n, 1))  # everybody 1|1: would change every count if it leaked in
    flank = [(10 + 7 * k, ["A", "C"], flank_gt) for k in range(flank_records)]
    spans = {}  # rid -> (first byte, end byte) of its records in the uncompressed stream
    for rid, recs in ((0, flank), (1, records), (2, flank)):
        begin = len(out)
        for pos, alleles, gt in recs:
            rlen = len(alleles[0])
            shared = struct.pack("<iiiIII", rid, pos, rlen, 0x7F800001, (len(alleles) << 16), (1 << 24) | n)
            shared += bytes([0x07])
            for a in alleles:
                shared += _typed_str(a)
            shared += bytes([0x00])
            indiv = bytes([0x11, 1, 0x21]) + np.asarray(gt, dtype=np.int8).tobytes()
            out += struct.pack("<II", len(shared), len(indiv)) + shared + indiv
        if recs:
            spans[rid] = (begin, len(out))
    member_off = []  # compressed offset of every member
    with open(path, "wb") as f:
        for i in range(0, len(out), member_bytes):  # several gzip members, like BGZF blocks
            chunk = bytes(out[i:i + member_bytes])
            member_off.append(f.tell())
            f.write(bgzf_block(chunk) if bgzf else gzip.compress(chunk, 6))
        member_off.append(f.tell())
        if bgzf:
            f.write(bgzf_block(b""))  # the EOF marker block
    if write_csi:
        assert bgzf

        def voff(byte):  # virtual offset of an uncompressed byte position
            k = byte // member_bytes
            return (member_off[k] << 16) | (byte - k * member_bytes)

        idx = bytearray(b"CSI\1" + struct.pack("<iii", 14, 5, 0) + struct.pack("<i", 3))
        pseudo = ((1 << 18) - 1) // 7 + 1
        for rid in range(3):
            if rid not in spans:
                idx += struct.pack("<i", 0)
                continue
            b, e = voff(spans[rid][0]), voff(spans[rid][1])
            idx += struct.pack("<i", 2)
            idx += struct.pack("<IQi", 0, b, 1) + struct.pack("<QQ", b, e)
            idx += struct.pack("<IQi", pseudo, 0, 2) + struct.pack("<QQ", b, e) + struct.pack("<QQ", len(records), 0)  # statistics, not a chunk
        idx += struct.pack("<Q", 0)
        with open(path + ".csi", "wb") as f:
            f.write(bgzf_block(bytes(idx)) + bgzf_block(b""))


def cohort_to_files(blk, pats, dirname, chrom="chrS", multiallelic_every=0, bgzf=False, member_bytes=40000, flank_records=0, write_csi=False):
    """Writes the file set of a synth.make_cohort block; returns the arguments of the reference's CLI."""
    m = blk.meta
    os.makedirs(dirname, exist_ok=True)
    fa = os.path.join(dirname, "genome.fa")
    write_fasta(fa, chrom, m["genome"].tobytes())
    beds = []
    for b, pm in enumerate(m["peak_map"]):
        p = os.path.join(dirname, "regions%d.bed" % (b + 1))
        write_bed(p, chrom, pm)
        beds.append(p)
    pwm = os.path.join(dirname, "pwms.txt")
    write_pwms(pwm, os.path.join(dirname, "thr"), pats)
    S = blk.n_samples
    samples = ["S%04d" % i for i in range(S)]
    recs = []
    allele = m["allele"].tobytes()
    bits = np.unpackbits(blk.carriers.view(np.uint8), axis=1, bitorder="little")[:, :2 * S]
    for v in range(len(m["var_pos"])):
        ref = allele[m["var_ref_off"][v]:m["var_ref_off"][v] + m["var_ref_len"][v]].decode()
        alt = allele[m["var_alt_off"][v]:m["var_alt_off"][v] + m["var_alt_len"][v]].decode()
        gt = np.empty((S, 2), dtype=np.int8)
        gt[:, 0] = np.where(bits[v, 0::2] == 1, 4, 2)
        gt[:, 1] = np.where(bits[v, 1::2] == 1, 5, 3)
        recs.append((int(m["var_pos"][v]), [ref, alt], gt))
        if multiallelic_every and v % multiallelic_every == 0:  # skipped by the reference (haplotype.rs:27,53-55)
            g2 = np.tile(np.array([[2, 7]], dtype=np.int8), (S, 1))
            recs.append((int(m["var_pos"][v]), [ref, "A", "C"], g2))
    write_bcf(os.path.join(dirname, "cohort.bcf"), chrom, m["genome_len"], samples, recs, member_bytes=member_bytes, bgzf=bgzf,
              flank_records=flank_records, write_csi=write_csi)
    names = [p["name"] for p in pats if p["direction"] == 0]
    return {"chromosome": chrom, "bcf": os.path.join(dirname, "cohort.bcf"), "beds": beds, "reference": fa, "pwm_file": pwm,
            "threshold_dir": os.path.join(dirname, "thr"), "names": names, "samples": samples}
