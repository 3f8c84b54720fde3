"""Shared helpers of the parity tests: run the same block through the CUDA path (C ABI) and through the CPU oracle."""
import ctypes as C

import numpy as np

from find_tfbs_b200 import binding
from find_tfbs_b200.binding import Block, INNER_DTYPE, VARIANT_DTYPE
from oracle import pyoracle as ora


def run_oracle(pattern_set, block, rows_mode=0, want_matches=False, n_threads=4, chunk=50):
    """The oracle takes the very same C structs (layouts are asserted equal in test_abi.py)."""
    cp = C.cast(pattern_set.c, C.POINTER(ora.TfbsPattern))
    cb = C.cast(C.pointer(block.c), C.POINTER(ora.TfbsBlock)).contents
    return ora.process_block(cp, pattern_set.n, cb, block.n_samples, rows_mode, want_matches, n_threads, chunk)


def run_gpu(pattern_set, block, rows_mode=0, record_matches=False, options=None, resident=False):
    ctx = binding.Context(0)
    try:
        ctx.set_option("rows_mode", rows_mode)
        ctx.set_option("record_matches", 1 if record_matches else 0)
        for k, v in (options or {}).items():
            ctx.set_option(k, v)
        ctx.set_patterns(pattern_set)
        if resident:
            ctx.upload_block(block)
            ctx.run_resident()
        else:
            ctx.submit_block(block)
        rows = ctx.collect()
        rows["stats"] = ctx.stats()
        if record_matches:
            rows["matches"] = ctx.matches(block.n_regions)
        return rows
    finally:
        ctx.close()


def assert_rows_equal(g, o):
    assert len(g["region"]) == len(o["region"]), "row count: gpu %d oracle %d" % (len(g["region"]), len(o["region"]))
    for k in ("region", "inner", "pattern_id", "vmin", "vmax", "left", "right"):
        assert np.array_equal(g[k], o[k]), "rows differ in " + k


def sorted_matches(region, pattern_index, group, start):
    m = np.stack([region.astype(np.int64), pattern_index.astype(np.int64), group.astype(np.int64), start.astype(np.int64)], axis=1)
    if len(m):
        m = m[np.lexsort((m[:, 3], m[:, 2], m[:, 1], m[:, 0]))]
    return m


def group_leaders(hap_group, n_regions):
    """Group numbers are labels (0 = reference); what is comparable is the partition.  Returns per haplotype the smallest
    haplotype of its group (-1 for the reference group) and a dict (region, group) -> that leader."""
    hg = hap_group.reshape(n_regions, -1).astype(np.int64)
    lead = np.full(hg.shape, -1, dtype=np.int64)
    table = {}
    for r in range(n_regions):
        for h, g in enumerate(hg[r]):
            if g == 0:
                continue
            if (r, g) not in table:
                table[(r, g)] = h
            lead[r, h] = table[(r, g)]
    return lead, table


def assert_matches_equal(g, o):
    gm = g["matches"]
    assert not gm["truncated"]
    n_regions = int(max(gm["region"].astype(np.int64).max(initial=-1), o["m_region"].astype(np.int64).max(initial=-1))) + 1
    H = 2 * g["left"].shape[1] if g["left"].ndim == 2 and g["left"].shape[1] else 0
    n_regions = len(o["hap_group"]) // H if H else n_regions
    gl, gt = group_leaders(gm["hap_group"], n_regions) if H else (np.zeros(0), {})
    ol, ot = group_leaders(o["hap_group"], n_regions) if H else (np.zeros(0), {})
    assert np.array_equal(gl, ol), "haplotype -> group partitions differ"
    ga = np.array([gt.get((int(r), int(x)), -1) for r, x in zip(gm["region"], gm["group"])], dtype=np.int64)
    oa = np.array([ot.get((int(r), int(x)), -1) for r, x in zip(o["m_region"], o["m_group"])], dtype=np.int64)
    a = sorted_matches(gm["region"], gm["pattern_index"], ga, gm["start"])
    b = sorted_matches(o["m_region"], o["m_pattern_index"], oa, o["m_start"])
    assert a.shape == b.shape, "hit count: gpu %d oracle %d" % (len(a), len(b))
    assert np.array_equal(a, b), "hit lists differ"


def check_stats(st, o):
    assert st["executed_cells"] == o["executed_cells"]
    assert st["nominal_cells"] == o["nominal_cells"]
    assert st["n_hits"] == o["n_hits"]
    assert st["n_groups"] == o["n_groups"]


def check_parity(pattern_set, block, rows_mode=0, options=None, matches=True, resident=False):
    """Oracle vs the CUDA path in both scan modes: full scan of every distinct haplotype (with the hit list when `matches`)
    and delta scoring (the default; patched haplotypes are scored only where a window touches a variant)."""
    o = run_oracle(pattern_set, block, rows_mode, matches)
    full = dict(options or {})
    full["delta"] = 0
    g = run_gpu(pattern_set, block, rows_mode, matches, full, resident)
    assert_rows_equal(g, o)
    if matches:
        assert_matches_equal(g, o)
    check_stats(g["stats"], o)
    assert g["stats"]["evaluated_cells"] == o["executed_cells"]
    d = run_gpu(pattern_set, block, rows_mode, False, options, resident)
    assert_rows_equal(d, o)
    check_stats(d["stats"], o)
    return g, o


def hand_block(n_samples, regions, variants, carriers_by_variant, inner=None):
    """Small explicit blocks.  regions: [(start, end, 'REFWINDOW')]; variants: [(region, pos, 'REF', 'ALT')] in record order;
    carriers_by_variant: list of haplotype index lists; inner: per region list of (start, end, bed, multiplicity)."""
    H = 2 * n_samples
    pitch = max(1, (H + 31) // 32)
    rs = np.array([r[0] for r in regions], dtype=np.int64)
    re = np.array([r[1] for r in regions], dtype=np.int64)
    ref = b"".join(r[2].encode() for r in regions)
    ref_off = np.cumsum([0] + [len(r[2]) for r in regions]).astype(np.uint64)
    var = np.zeros(len(variants), dtype=VARIANT_DTYPE)
    allele = bytearray()
    var_off = np.zeros(len(regions) + 1, dtype=np.uint32)
    car = np.zeros((max(1, len(variants)), pitch), dtype=np.uint32)
    order = sorted(range(len(variants)), key=lambda i: variants[i][0])  # stable: keeps record order inside a region
    for k, i in enumerate(order):
        r, pos, rf, al = variants[i]
        var[k]["pos"] = pos
        var[k]["ref_off"] = len(allele)
        var[k]["ref_len"] = len(rf)
        allele += rf.encode()
        var[k]["alt_off"] = len(allele)
        var[k]["alt_len"] = len(al)
        allele += al.encode()
        var[k]["carrier_row"] = k
        for h in carriers_by_variant[i]:
            car[k, h // 32] |= np.uint32(1 << (h % 32))
        var_off[r + 1] += 1
    var_off = np.cumsum(var_off).astype(np.uint32)
    if inner is None:
        inner = [[(int(rs[i]), int(re[i]), 0, 1)] for i in range(len(regions))]
    inn = np.zeros(sum(len(x) for x in inner), dtype=INNER_DTYPE)
    inner_off = np.zeros(len(regions) + 1, dtype=np.uint32)
    k = 0
    for r, lst in enumerate(inner):
        for (s, e, b, m) in lst:
            inn[k] = (s, e, b, m)
            k += 1
        inner_off[r + 1] = k
    return Block(n_samples, rs, re, ref_off, np.frombuffer(ref, dtype=np.uint8) if ref else np.zeros(0, np.uint8), inner_off, inn, var_off, var,
                 np.frombuffer(bytes(allele), dtype=np.uint8) if allele else np.zeros(0, np.uint8), car)


def lowered(pattern_set):
    """The same patterns with every threshold lowered by one: hits of that list = hits + windows scoring exactly min_score."""
    return binding.PatternSet([dict(p, min_score=int(p.get("min_score", 0)) - 1) for p in pattern_set.items])


def oracle_audit(pattern_set, block):
    """Ties (score == min_score) as the difference of two oracle hit lists, keyed by (region, pattern, group leader, start);
    per-haplotype flags from the oracle's load_haplotypes."""
    a = run_oracle(lowered(pattern_set), block, 1, True)
    b = run_oracle(pattern_set, block, 1, True)
    assert np.array_equal(a["hap_group"], b["hap_group"])
    H = 2 * block.n_samples
    _, table = group_leaders(b["hap_group"], block.n_regions) if H else (None, {})

    def keyed(o):
        return set((int(r), int(p), table.get((int(r), int(g)), -1), int(s)) for r, p, g, s in zip(o["m_region"], o["m_pattern_index"], o["m_group"], o["m_start"]))

    return {"ties": keyed(a) - keyed(b), "hap_flags": b["hap_flags"], "n_hits": b["n_hits"]}


def gpu_audit(pattern_set, block, options=None):
    ctx = binding.Context(0)
    try:
        for k, v in (options or {}).items():
            ctx.set_option(k, v)
        ctx.set_patterns(pattern_set)
        ctx.upload_block(block)
        au = ctx.audit()
        rows = ctx.collect()  # the audit leaves a normal run of the block behind
        st = ctx.stats()
    finally:
        ctx.close()
    H = 2 * block.n_samples
    _, table = group_leaders(au["hap_group"], block.n_regions) if H else (None, {})
    au["ties"] = set((int(r), int(p), table.get((int(r), int(g)), -1), int(s)) for r, p, g, s in zip(au["region"], au["pattern_index"], au["group"], au["start"]))
    return au, rows, st
