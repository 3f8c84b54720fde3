"""CPU tests of the C++ driver's host logic (find_tfbs_b200/csrc/driver/driver.cpp compiled with a test shim, no GPU, no main):
weight / threshold / PWM parsing, range merging, BCF and FASTA decoding and the row finaliser, each against the oracle or the
reference's fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from find_tfbs_b200 import synth
from oracle import pyoracle as ora
import file_writers as fw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def drv(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("drv") / "libdrvhost.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTFBS_DRIVER_NO_MAIN", "-DTFBS_DRIVER_TEST_SHIM",
                           "-static-libstdc++", "-static-libgcc", "-Wl,-Bsymbolic", "-Wl,--exclude-libs,ALL", "-o", so,
                           os.path.join(ROOT, "find_tfbs_b200", "csrc", "driver", "driver.cpp"), "-lz"])
    lib = C.CDLL(so)
    lib.drv_last_error.restype = C.c_char_p
    return lib


def test_parse_weight_matches_oracle(drv):
    rng = np.random.default_rng(0)
    cases = ["1.0", "3.999", "-4.4", "0.0625", "-0.0625", "0.0005", "2.5e-3", "-28.912716067144597", "0", "1e-9", "7.7515"]
    cases += ["%.*f" % (int(rng.integers(1, 9)), x) for x in rng.normal(0, 3, size=300)]
    for s in cases:
        out = C.c_int32()
        assert drv.drv_parse_weight(s.encode(), C.byref(out)) == 0
        assert out.value == ora.parse_weight(s), s
    assert drv.drv_parse_weight(b"abc", C.byref(C.c_int32())) == -1


def test_threshold_and_pwm_parsing(drv, golden_dir, tmp_path):
    for thr in (1e-4, 1e-3, 0.5, 1e-6, 2.0):
        out = C.c_int32()
        found = drv.drv_parse_threshold_file(os.path.join(golden_dir, "ACGT.thr").encode(), C.c_float(thr), C.byref(out))
        exp = ora.parse_threshold_file(os.path.join(golden_dir, "ACGT.thr"), thr)
        assert (out.value if found else None) == exp
    pats = synth.make_pwms(7, seed=4, lmin=5, lmax=20)
    pwm = str(tmp_path / "pwms.txt")
    fw.write_pwms(pwm, str(tmp_path / "thr"), pats)
    names = [p["name"] for p in pats if p["direction"] == 0]
    wanted = [names[3], names[0], "MISSING", names[5]]  # pattern_id follows the FILE order of the wanted PWMs (pattern.rs:61,69,81)
    for fwd_only in (0, 1):
        cap = 64
        lens = (C.c_uint32 * cap)()
        ms = (C.c_int32 * cap)()
        pid = (C.c_uint16 * cap)()
        dr = (C.c_uint8 * cap)()
        w = (C.c_int32 * 8192)()
        # the threshold file of MISSING does not exist: the reference panics in read_lines (pattern.rs:115)
        assert drv.drv_parse_pwms(pwm.encode(), str(tmp_path / "thr").encode(), C.c_float(1e-4), ",".join(wanted).encode(), fwd_only,
                                  cap, lens, ms, pid, dr, w, 8192) == -1
        assert b"Could not open file" in drv.drv_last_error()
        ok = [x for x in wanted if x != "MISSING"]
        n = drv.drv_parse_pwms(pwm.encode(), str(tmp_path / "thr").encode(), C.c_float(1e-4), ",".join(ok).encode(), fwd_only, cap, lens, ms, pid,
                               dr, w, 8192)
        exp = ora.parse_pwm_files(pwm, str(tmp_path / "thr"), 1e-4, ok, not fwd_only)
        assert n == len(exp)
        off = 0
        for i, p in enumerate(exp):
            L = p["weights"].shape[0]
            assert (lens[i], ms[i], pid[i], dr[i]) == (L, p["min_score"], p["pattern_id"], p["direction"])
            assert [w[off + k] for k in range(4 * L)] == p["weights"].reshape(-1).tolist()
            off += 4 * L


def test_merge_ranges_matches_oracle(drv, tmp_path):
    rng = np.random.default_rng(1)
    for _ in range(40):
        n = int(rng.integers(1, 40))
        s = rng.integers(0, 500, size=n).astype(np.uint64)
        e = (s + rng.integers(0, 60, size=n).astype(np.uint64)).astype(np.uint64)
        os_ = (C.c_uint64 * n)()
        oe = (C.c_uint64 * n)()
        m = drv.drv_merge_ranges(s.ctypes.data_as(C.POINTER(C.c_uint64)), e.ctypes.data_as(C.POINTER(C.c_uint64)), n, os_, oe)
        bed = str(tmp_path / "r.bed")
        with open(bed, "w") as f:
            for a, b in zip(s, e):
                f.write("chrT\t%d\t%d\n" % (a, b))
        merged, _ = ora.load_peak_files([bed], "chrT", 0)
        assert [(os_[i], oe[i]) for i in range(m)] == merged


def test_finalise_row_matches_oracle(drv):
    rng = np.random.default_rng(2)
    for trial in range(300):
        S = int(rng.integers(1, 40))
        hi = int(rng.choice([1, 2, 3, 8, 40, 1000]))
        l = rng.integers(0, hi + 1, size=S).astype(np.uint32)
        r = rng.integers(0, hi + 1, size=S).astype(np.uint32)
        if trial % 7 == 0:
            r = (hi - l).astype(np.uint32)  # constant sum: the row is dropped (main.rs:456-458)
        min_maf = int(rng.integers(0, 4))
        info = C.create_string_buffer(4096)
        gt = C.create_string_buffer(64 * S + 64)
        keep = drv.drv_finalise_row(l.ctypes.data_as(C.POINTER(C.c_uint32)), r.ctypes.data_as(C.POINTER(C.c_uint32)), S, min_maf, info, 4096, gt, len(gt))
        exp = ora.counts_as_genotypes(l, r)
        if exp is None or exp["maf"] < min_maf:
            assert keep == 0
            continue
        assert keep == 1
        assert info.value.decode() == "COUNTS=%s;freqs=%d/%d/%d" % (",".join(str(c) for c in exp["counts"]), *exp["freqs"])
        assert gt.value.decode() == exp["genotypes"]


def test_bcf_and_fasta_decoding(drv, golden_dir, tmp_path):
    def load(bcf, samples, chrom, cap=100000, ccap=4000000, use_index=1):
        pos = (C.c_int64 * cap)()
        na = (C.c_uint32 * cap)()
        row = (C.c_uint32 * cap)()
        car = (C.c_uint32 * ccap)()
        n = C.c_uint32()
        ns = C.c_uint32()
        pitch = C.c_uint32()
        rc = drv.drv_load_bcf(bcf.encode(), (samples or "").encode(), chrom.encode(), cap, pos, na, row, car, ccap, C.byref(n), C.byref(ns), C.byref(pitch), use_index)
        assert rc == 0, drv.drv_last_error()
        return [pos[i] for i in range(n.value)], [na[i] for i in range(n.value)], [row[i] for i in range(n.value)], car, ns.value, pitch.value

    # the reference's fixture: one record at 100, INDIVIDUAL1 = 1|0
    pos, na, row, car, ns, pitch = load(os.path.join(golden_dir, "genotypes2.bcf"), os.path.join(golden_dir, "samples"), "chr1")
    assert (pos, na, row, ns, pitch) == ([100], [2], [0], 4, 1) and car[0] == 1
    # a synthetic cohort written by the test writers: carrier bits equal the generator's, multi-allelic records get no row
    pats = synth.make_pwms(2, seed=5, lmin=6, lmax=10)
    blk = synth.make_cohort(21, 12, seed=5, lmax_pattern=10, region_len=(50, 150), variant_rate=0.05)
    a = fw.cohort_to_files(blk, pats, str(tmp_path / "c"), multiallelic_every=5)
    b = fw.cohort_to_files(blk, pats, str(tmp_path / "b"), multiallelic_every=5, bgzf=True, member_bytes=1500)  # dozens of real BGZF members: parallel inflate
    assert ora.read_bcf(b["bcf"])["pos"] == ora.read_bcf(a["bcf"])["pos"]
    assert load(b["bcf"], None, b["chromosome"])[:3] == load(a["bcf"], None, a["chromosome"])[:3]
    pos, na, row, car, ns, pitch = load(a["bcf"], None, a["chromosome"])
    o = ora.read_bcf(a["bcf"])
    assert pos == o["pos"] and na == [len(x) for x in o["alleles"]] and ns == 21 and pitch == blk.carriers.shape[1]
    rows = [r for r in row if r != 0xFFFFFFFF]
    assert rows == list(range(len(blk.meta["var_pos"])))
    got = np.array([car[i] for i in range(len(rows) * pitch)], dtype=np.uint32).reshape(len(rows), pitch)
    assert np.array_equal(got, blk.carriers[:len(rows)])
    # a multi-contig BGZF file with a CSI index: only the members of the wanted contig are read, the records are the same as
    # without the index (whole-file scan) and as in the single-contig file; the other contigs' records never show up
    for mb in (1500, 700, 613, 631, 653, 777, 100000):  # the contig's last member usually ends inside a record of the next contig
        c = fw.cohort_to_files(blk, pats, str(tmp_path / ("i%d" % mb)), multiallelic_every=5, bgzf=True, member_bytes=mb, flank_records=40, write_csi=True)
        assert os.path.exists(c["bcf"] + ".csi")
        with_index = load(c["bcf"], None, c["chromosome"])
        assert with_index[:3] == load(c["bcf"], None, c["chromosome"], use_index=0)[:3] == load(a["bcf"], None, a["chromosome"])[:3]
        assert [with_index[3][i] for i in range(len(rows) * pitch)] == [car[i] for i in range(len(rows) * pitch)]
        assert len(load(c["bcf"], None, "chrOther")[0]) == 40 and len(load(c["bcf"], None, "chrZ")[0]) == 40
        assert load(c["bcf"], None, "chrZ")[0] == load(c["bcf"], None, "chrZ", use_index=0)[0]
    # the reference's own index (written by bcftools): same single record with and without it
    assert load(os.path.join(golden_dir, "genotypes2.bcf"), None, "chr1")[:3] == load(os.path.join(golden_dir, "genotypes2.bcf"), None, "chr1", use_index=0)[:3] == ([100], [2], [0])
    # sample subset: columns follow the BCF order whatever the file order is (main.rs:293-313)
    sub = str(tmp_path / "subset")
    open(sub, "w").write("\n".join([a["samples"][9], a["samples"][2], "x", a["samples"][17]]) + "\n")
    pos2, _, row2, car2, ns2, pitch2 = load(a["bcf"], sub, a["chromosome"])
    assert ns2 == 3 and pitch2 == 1
    bits = np.unpackbits(blk.carriers.view(np.uint8), axis=1, bitorder="little")
    want = bits[:len(rows), [4, 5, 18, 19, 34, 35]]
    got2 = np.array([[(car2[i] >> k) & 1 for k in range(6)] for i in range(len(rows))], dtype=np.uint8)
    assert np.array_equal(got2, want)
    # FASTA slices through the .fai; an interval that ends behind the contig is an error, as in bio's IndexedReader (main.rs:157-159)
    g = blk.meta["genome"].tobytes()
    for start, stop in ((0, 10), (57, 63), (59, 61), (60, 200), (len(g) - 5, len(g)), (123, 123)):
        buf = (C.c_uint8 * 4096)()
        n = C.c_uint64()
        assert drv.drv_fasta_fetch(a["reference"].encode(), a["chromosome"].encode(), start, stop, buf, 4096, C.byref(n)) == 0
        assert bytes(buf[:n.value]) == g[start:stop]
    assert drv.drv_fasta_fetch(a["reference"].encode(), a["chromosome"].encode(), len(g) - 5, len(g) + 50, (C.c_uint8 * 4096)(), 4096, C.byref(C.c_uint64())) == -1
    assert drv.drv_fasta_fetch(a["reference"].encode(), b"chrNope", 0, 5, (C.c_uint8 * 16)(), 16, C.byref(C.c_uint64())) == -1


def test_bgzf_writer(drv, tmp_path):
    """The VCF writer (main.rs:264-290 uses bgzip::BGzWriter): valid BGZF members (BC subfield, BSIZE, CRC, ISIZE, EOF block) whatever
    the write granularity and thread count; the concatenation inflates to the input; the driver's own parallel reader reads it back."""
    import struct
    import zlib
    rng = np.random.default_rng(3)
    text = "\n".join("1\t%d\tregions.bed,PWM%d,%d-%d\t.\t.\t.\tPASS\tCOUNTS=0,1;freqs=5/0/1\tGT:DS" % (i, i % 7, i * 10, i * 10 + 300) +
                     "".join(rng.choice(["\t0|0:0.0", "\t1|1:2.0", "\t0|1:0.6667"], size=40)) for i in range(4000)).encode()
    for threads, piece in ((1, 1 << 20), (4, 777), (3, 200000)):
        path = str(tmp_path / ("t%d.vcf.gz" % threads))
        assert drv.drv_write_bgzf(path.encode(), text, len(text), threads, piece) == 0
        raw = open(path, "rb").read()
        off, out, n_members = 0, b"", 0
        while off < len(raw):
            assert raw[off:off + 4] == b"\x1f\x8b\x08\x04" and raw[off + 12:off + 16] == b"BC\x02\x00"
            bsize = struct.unpack("<H", raw[off + 16:off + 18])[0] + 1
            crc, isize = struct.unpack("<II", raw[off + bsize - 8:off + bsize])
            data = zlib.decompress(raw[off + 18:off + bsize - 8], -15)
            assert len(data) == isize <= 0xff00 and zlib.crc32(data) & 0xFFFFFFFF == crc
            out += data
            off += bsize
            n_members += 1
        assert out == text and n_members > 5
        assert raw.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))  # the BGZF EOF block
        assert ora.gunzip_file(path).encode() == text


def test_threshold_tooling_is_exact(drv, tmp_path):
    """score_threshold / write_threshold_file of the driver (--write_thresholds): the exact DP equals the harness's
    (synth.score_threshold), and a written .thr file read back with the reference's rule (pattern.rs:18-35, the last line whose
    p-value exceeds the threshold) gives the same score for every listed p-value."""
    pats = [p for p in synth.make_pwms(12, seed=8, lmin=5, lmax=30) if p["direction"] == 0]
    pvals = [1e-2, 1e-3, 5e-4, 1e-4, 1e-5]
    for i, p in enumerate(pats):
        w = np.ascontiguousarray(p["weights"], dtype=np.int32)
        for pv in pvals + [0.3, 1e-7]:
            out = C.c_int32()
            assert drv.drv_score_threshold(w.ctypes.data_as(C.POINTER(C.c_int32)), w.shape[0], C.c_double(pv), C.byref(out)) == 0
            assert out.value == synth.score_threshold(w, pv), (i, pv)
        path = str(tmp_path / ("P%d.thr" % i))
        arr = (C.c_double * len(pvals))(*pvals)
        assert drv.drv_write_threshold_file(path.encode(), w.ctypes.data_as(C.POINTER(C.c_int32)), w.shape[0], arr, len(pvals)) == 0
        for pv in pvals:
            assert ora.parse_threshold_file(path, pv) == synth.score_threshold(w, pv), (i, pv)
