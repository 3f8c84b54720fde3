"""ctypes view of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; nothing under find_tfbs_b200/ does.  The oracle restates the reference's algorithm
(see oracle.hpp for the file:line citations); this file only marshals arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only; no reference sources are used)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_io.cpp", "oracle_c.cpp", "oracle.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "tfbs.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ora_last_error.restype = C.c_char_p
    return _LIB


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("oracle error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().ora_last_error().decode())


# ---- C structs shared with include/tfbs.h (kept in sync with find_tfbs_b200/binding.py by a test) ----
class TfbsPattern(C.Structure):
    _fields_ = [("weights", C.POINTER(C.c_int32)), ("len", C.c_uint32), ("min_score", C.c_int32),
                ("pattern_id", C.c_uint16), ("direction", C.c_uint8), ("kind", C.c_uint8)]


class TfbsInnerRegion(C.Structure):
    _fields_ = [("start", C.c_int64), ("end", C.c_int64), ("bed_index", C.c_uint32), ("multiplicity", C.c_uint32)]


class TfbsVariant(C.Structure):
    _fields_ = [("pos", C.c_int64), ("ref_off", C.c_uint32), ("ref_len", C.c_uint32), ("alt_off", C.c_uint32),
                ("alt_len", C.c_uint32), ("carrier_row", C.c_uint32), ("reserved", C.c_uint32)]


class TfbsBlock(C.Structure):
    _fields_ = [("n_regions", C.c_uint32), ("n_samples", C.c_uint32),
                ("region_start", C.POINTER(C.c_int64)), ("region_end", C.POINTER(C.c_int64)),
                ("ref_off", C.POINTER(C.c_uint64)), ("ref_bases", C.POINTER(C.c_uint8)),
                ("inner_off", C.POINTER(C.c_uint32)), ("inner", C.POINTER(TfbsInnerRegion)),
                ("var_off", C.POINTER(C.c_uint32)), ("variants", C.POINTER(TfbsVariant)),
                ("allele_bases", C.POINTER(C.c_uint8)), ("allele_bytes", C.c_uint64),
                ("carriers", C.POINTER(C.c_uint32)), ("n_carrier_rows", C.c_uint32), ("carrier_pitch", C.c_uint32)]


class _BlockResult(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("region", C.POINTER(C.c_uint32)), ("inner", C.POINTER(C.c_uint32)),
                ("pattern_id", C.POINTER(C.c_uint16)), ("vmin", C.POINTER(C.c_uint32)), ("vmax", C.POINTER(C.c_uint32)),
                ("left", C.POINTER(C.c_uint32)), ("right", C.POINTER(C.c_uint32)),
                ("n_matches", C.c_uint64), ("m_region", C.POINTER(C.c_uint32)), ("m_pattern_index", C.POINTER(C.c_uint32)),
                ("m_group", C.POINTER(C.c_uint32)), ("m_start", C.POINTER(C.c_int64)),
                ("n_hap_group", C.c_uint64), ("hap_group", C.POINTER(C.c_uint32)),
                ("executed_cells", C.c_uint64), ("nominal_cells", C.c_uint64), ("n_groups", C.c_uint64), ("n_hits", C.c_uint64),
                ("collision_regions", C.c_uint32), ("truncated_regions", C.c_uint32),
                ("n_hap_flags", C.c_uint64), ("hap_flags", C.POINTER(C.c_uint8))]


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def process_block(c_patterns, n_patterns, c_block, n_samples, rows_mode=0, want_matches=False, n_threads=1, chunk=50):
    """c_patterns: (TfbsPattern * n) array; c_block: TfbsBlock.  Returns a dict of numpy arrays.
    chunk: regions per work-queue item (50 in the reference, main.rs:378)."""
    res = _BlockResult()
    lib().ora_set_chunk_size(C.c_uint32(chunk))
    _check(lib().ora_process_block(c_patterns, C.c_uint32(n_patterns), C.byref(c_block), C.c_int(rows_mode),
                                   C.c_int(1 if want_matches else 0), C.c_int(n_threads), C.byref(res)))
    n = res.n_rows
    out = {
        "region": _arr(res.region, n, np.uint32), "inner": _arr(res.inner, n, np.uint32),
        "pattern_id": _arr(res.pattern_id, n, np.uint16), "vmin": _arr(res.vmin, n, np.uint32),
        "vmax": _arr(res.vmax, n, np.uint32),
        "left": _arr(res.left, n * n_samples, np.uint32).reshape(n, n_samples),
        "right": _arr(res.right, n * n_samples, np.uint32).reshape(n, n_samples),
        "m_region": _arr(res.m_region, res.n_matches, np.uint32),
        "m_pattern_index": _arr(res.m_pattern_index, res.n_matches, np.uint32),
        "m_group": _arr(res.m_group, res.n_matches, np.uint32),
        "m_start": _arr(res.m_start, res.n_matches, np.int64),
        "hap_group": _arr(res.hap_group, res.n_hap_group, np.uint32),
        "hap_flags": _arr(res.hap_flags, res.n_hap_flags, np.uint8),
        "executed_cells": int(res.executed_cells), "nominal_cells": int(res.nominal_cells),
        "n_groups": int(res.n_groups), "n_hits": int(res.n_hits),
        "collision_regions": int(res.collision_regions), "truncated_regions": int(res.truncated_regions),
    }
    lib().ora_free_block_result(C.byref(res))
    return out


_CODE = "ACGTN"


def patch_haplotype(rng, diffs, ref):
    """rng=(start,end); diffs=[(pos, 'REF', 'ALT')]; ref=[('A', pos), ...] -> ([(letter, pos)], truncated)."""
    n = len(diffs)
    pos = (C.c_uint64 * max(1, n))(*[d[0] for d in diffs])
    refs = (C.c_char_p * max(1, n))(*[d[1].encode() for d in diffs])
    alts = (C.c_char_p * max(1, n))(*[d[2].encode() for d in diffs])
    letters = "".join(r[0] for r in ref).encode()
    rpos = (C.c_uint64 * max(1, len(ref)))(*[r[1] for r in ref])
    cap = len(ref) + sum(len(d[2]) for d in diffs) + 8
    out_nuc = (C.c_uint8 * cap)()
    out_pos = (C.c_uint64 * cap)()
    out_len = C.c_uint32()
    trunc = C.c_int()
    _check(lib().ora_patch_haplotype(C.c_uint64(rng[0]), C.c_uint64(rng[1]), C.c_uint32(n), pos, refs, alts, C.c_uint32(len(ref)),
                                     letters, rpos, C.c_uint32(cap), out_nuc, out_pos, C.byref(out_len), C.byref(trunc)))
    return [(_CODE[out_nuc[i]], out_pos[i]) for i in range(out_len.value)], bool(trunc.value)


def matches(weights, min_score, pattern_id, hap):
    """weights: [[a,c,g,t],...]; hap=[('A',pos),...] -> [(start,end,pattern_id)]."""
    w = np.ascontiguousarray(np.array(weights, dtype=np.int32).reshape(-1, 4))
    letters = "".join(h[0] for h in hap).encode()
    pos = (C.c_uint64 * max(1, len(hap)))(*[h[1] for h in hap])
    cap = len(hap) + 1
    s = (C.c_uint64 * cap)()
    e = (C.c_uint64 * cap)()
    p = (C.c_uint16 * cap)()
    n = C.c_uint32()
    _check(lib().ora_matches(w.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint32(w.shape[0]), C.c_int32(min_score), C.c_uint16(pattern_id),
                             C.c_uint32(len(hap)), letters, pos, C.c_uint32(cap), s, e, p, C.byref(n)))
    return [(s[i], e[i], p[i]) for i in range(n.value)]


def parse_weight(s):
    out = C.c_int32()
    _check(lib().ora_parse_weight(s.encode(), C.byref(out)))
    return out.value


def reverse_complement(weights):
    w = np.ascontiguousarray(np.array(weights, dtype=np.int32).reshape(-1, 4))
    out = np.zeros_like(w)
    lib().ora_reverse_complement(w.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint32(w.shape[0]), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def parse_threshold_file(path, threshold):
    out = C.c_int32()
    found = C.c_int()
    _check(lib().ora_parse_threshold_file(path.encode(), C.c_float(threshold), C.byref(out), C.byref(found)))
    return out.value if found.value else None


def _take_str(p):
    s = C.cast(p, C.c_char_p).value.decode()
    lib().ora_free(p)
    return s


def parse_pwm_files(pwm_file, threshold_dir, threshold, wanted, add_reverse=True):
    n = C.c_uint32()
    lens = C.POINTER(C.c_uint32)()
    ws = C.POINTER(C.c_int32)()
    pid = C.POINTER(C.c_uint16)()
    ms = C.POINTER(C.c_int32)()
    dr = C.POINTER(C.c_uint8)()
    names = C.c_void_p()
    _check(lib().ora_parse_pwm_files(pwm_file.encode(), threshold_dir.encode(), C.c_float(threshold), ",".join(wanted).encode(),
                                     C.c_int(1 if add_reverse else 0), C.byref(n), C.byref(lens), C.byref(ws), C.byref(pid),
                                     C.byref(ms), C.byref(dr), C.byref(names)))
    out = []
    nm = _take_str(names).split("\n")
    off = 0
    for i in range(n.value):
        L = lens[i]
        w = np.array([ws[off + k] for k in range(4 * L)], dtype=np.int32).reshape(L, 4)
        off += 4 * L
        out.append({"name": nm[i], "weights": w, "pattern_id": pid[i], "min_score": ms[i], "direction": dr[i]})
    for p in (lens, ws, pid, ms, dr):
        lib().ora_free(p)
    return out


def count_matches(match_list, inner_peaks, sample_count):
    """match_list=[(start,end,pattern_id,sample,side)], inner_peaks=[(bed,start,end)] -> {(bed,start,end,pid): (left,right)}."""
    n = len(match_list)
    ms = (C.c_uint64 * max(1, n))(*[m[0] for m in match_list])
    me = (C.c_uint64 * max(1, n))(*[m[1] for m in match_list])
    mp = (C.c_uint16 * max(1, n))(*[m[2] for m in match_list])
    mh = (C.c_uint32 * max(1, n))(*[2 * m[3] + m[4] for m in match_list])
    k = len(inner_peaks)
    ib = (C.c_uint32 * max(1, k))(*[p[0] for p in inner_peaks])
    is_ = (C.c_uint64 * max(1, k))(*[p[1] for p in inner_peaks])
    ie = (C.c_uint64 * max(1, k))(*[p[2] for p in inner_peaks])
    cap = max(1, n * max(1, k))
    ob = (C.c_uint32 * cap)()
    os_ = (C.c_uint64 * cap)()
    oe = (C.c_uint64 * cap)()
    op = (C.c_uint16 * cap)()
    ol = (C.c_uint32 * (cap * sample_count))()
    orr = (C.c_uint32 * (cap * sample_count))()
    on = C.c_uint32()
    lib().ora_count_matches(C.c_uint32(n), ms, me, mp, mh, C.c_uint32(k), ib, is_, ie, C.c_uint32(sample_count), C.c_uint32(cap),
                            ob, os_, oe, op, ol, orr, C.byref(on))
    out = {}
    for i in range(on.value):
        out[(ob[i], os_[i], oe[i], op[i])] = ([ol[i * sample_count + s] for s in range(sample_count)],
                                              [orr[i * sample_count + s] for s in range(sample_count)])
    return out


def counts_as_genotypes(v1, v2):
    a = np.ascontiguousarray(np.array(v1, dtype=np.uint32))
    b = np.ascontiguousarray(np.array(v2, dtype=np.uint32))
    maf = C.c_uint32()
    freqs = (C.c_uint32 * 3)()
    cc = C.c_void_p()
    gg = C.c_void_p()
    rc = lib().ora_counts_as_genotypes(a.ctypes.data_as(C.POINTER(C.c_uint32)), b.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(len(a)),
                                       C.byref(maf), freqs, C.byref(cc), C.byref(gg))
    if rc == 0:
        return None
    return {"counts": [int(x) for x in _take_str(cc).split(",")], "maf": maf.value, "freqs": (freqs[0], freqs[1], freqs[2]),
            "genotypes": _take_str(gg)}


def load_peak_files(beds, chromosome, after_position=0):
    nm = C.c_uint32()
    ms = C.POINTER(C.c_uint64)()
    me = C.POINTER(C.c_uint64)()
    nf = C.c_uint32()
    off = C.POINTER(C.c_uint32)()
    fs = C.POINTER(C.c_uint64)()
    fe = C.POINTER(C.c_uint64)()
    names = C.c_void_p()
    _check(lib().ora_load_peak_files(",".join(beds).encode(), chromosome.encode(), C.c_uint64(after_position), C.byref(nm), C.byref(ms),
                                     C.byref(me), C.byref(nf), C.byref(off), C.byref(fs), C.byref(fe), C.byref(names)))
    merged = [(ms[i], me[i]) for i in range(nm.value)]
    nmz = _take_str(names).split("\n")
    peak_map = {}
    for b in range(nf.value):
        peak_map[nmz[b]] = [(fs[k], fe[k]) for k in range(off[b], off[b + 1])]
    for p in (ms, me, off, fs, fe):
        lib().ora_free(p)
    return merged, peak_map


def range_overlaps(a, b):
    return bool(lib().ora_range_overlaps(C.c_uint64(a[0]), C.c_uint64(a[1]), C.c_uint64(b[0]), C.c_uint64(b[1])))


def range_contains(a, p):
    return bool(lib().ora_range_contains(C.c_uint64(a[0]), C.c_uint64(a[1]), C.c_uint64(p)))


def run(chromosome, bcf, beds, reference, samples_file, pwm_file, threshold_dir, threshold, wanted, output="",
        forward_only=False, min_maf=0, threads=1, after_position=0):
    text = C.c_void_p()
    _check(lib().ora_run(chromosome.encode(), bcf.encode(), ",".join(beds).encode(), reference.encode(),
                         (samples_file or "").encode(), pwm_file.encode(), threshold_dir.encode(), C.c_float(threshold),
                         ",".join(wanted).encode(), output.encode(), C.c_int(1 if forward_only else 0), C.c_uint32(min_maf),
                         C.c_uint32(threads), C.c_uint64(after_position), C.byref(text)))
    return _take_str(text)


def gunzip_file(path):
    text = C.c_void_p()
    _check(lib().ora_gunzip_file(path.encode(), C.byref(text)))
    return _take_str(text)


def read_bcf(path):
    n = C.c_uint32()
    pos = C.POINTER(C.c_int64)()
    al = C.c_void_p()
    ns = C.c_uint32()
    gt = C.POINTER(C.c_int32)()
    sn = C.c_void_p()
    cn = C.c_void_p()
    _check(lib().ora_read_bcf(path.encode(), C.byref(n), C.byref(pos), C.byref(al), C.byref(ns), C.byref(gt), C.byref(sn), C.byref(cn)))
    out = {"pos": [pos[i] for i in range(n.value)], "alleles": [a.split(",") for a in _take_str(al).split("\n")[:-1]],
           "samples": _take_str(sn).split("\n")[:-1], "contigs": _take_str(cn).split("\n")[:-1],
           "gt": np.array([gt[i] for i in range(n.value * ns.value * 2)], dtype=np.int32).reshape(n.value, ns.value, 2)}
    lib().ora_free(pos)
    lib().ora_free(gt)
    return out
