"""Seeded synthetic cohorts for the find-tfbs hot path (BASELINE.md "Workloads").

Produces the arrays of a tfbs_block directly (what the reference's BCF / FASTA / BED loaders would hand to
process_peak, main.rs:395-436) plus a pattern list shaped like parse_pwm_files' output (pattern.rs:37-87).
Nothing here reads /root/reference or the oracle.
"""
import numpy as np

from .binding import Block, INNER_DTYPE, VARIANT_DTYPE, DIR_N, DIR_P, PATTERN_PWM

LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


# ---- reference semantics mirrored on the host side (tested against the oracle) -----------------------
def range_overlaps(a, b):
    """Range::overlaps (range.rs:18-21): b.start in a or b.end in a -- asymmetric on purpose."""
    return (a[0] <= b[0] <= a[1]) or (a[0] <= b[1] <= a[1])


def merge_regions(regions):
    """RangeStack (range.rs:43-87): stable sort by start, fold with last.overlaps(range)."""
    out = []
    for s, e in sorted(regions, key=lambda r: r[0]):
        if out and range_overlaps(out[-1], (s, e)):
            out[-1] = (min(out[-1][0], s), max(out[-1][1], e))
        else:
            out.append((s, e))
    return out


def select_inner_peaks(merged, peak_map):
    """select_inner_peaks (main.rs:62-72): p.overlaps(merged); identical occurrences collapse into a multiplicity."""
    out = []
    for b, peaks in enumerate(peak_map):
        seen = {}
        for p in peaks:
            if range_overlaps(p, merged):
                if p in seen:
                    out[seen[p]][3] += 1
                else:
                    seen[p] = len(out)
                    out.append([p[0], p[1], b, 1])
    return out


# ---- PWMs ------------------------------------------------------------------------------------------------
def score_threshold(w, pvalue):
    """Largest integer s with P(score >= s) > pvalue under uniform ACGT (what the last qualifying line of a HOCOMOCO
    .thr file holds, pattern.rs:18-35); hits are windows with score > s."""
    w = np.asarray(w, dtype=np.int64)
    lo = w.min(axis=1)
    span = int((w.max(axis=1) - lo).sum())
    dist = np.zeros(span + 1)
    dist[0] = 1.0
    top = 0
    for c in range(w.shape[0]):
        new = np.zeros(span + 1)
        for b in range(4):
            d = int(w[c, b] - lo[c])
            new[d:d + top + 1] += 0.25 * dist[:top + 1]
        top += int(w[c].max() - lo[c])
        dist = new
    tail = np.cumsum(dist[::-1])[::-1]  # tail[k] = P(score - sum(lo) >= k)
    ok = np.nonzero(tail > pvalue)[0]
    k = int(ok.max()) if len(ok) else 0
    return int(k + lo.sum())


def make_pwms(n_pwms, seed, lmin=8, lmax=30, pvalue=1e-4, both_strands=True):
    """Random log-odds PWMs x 1000 (same scale as pattern.rs:13-16), forward + reverse complement sharing pattern_id."""
    rng = np.random.default_rng(seed)
    pats = []
    for i in range(n_pwms):
        L = int(rng.integers(lmin, lmax + 1))
        alpha = rng.choice([0.08, 0.3, 1.0, 6.0], size=L, p=[0.25, 0.3, 0.25, 0.2])
        p = np.stack([rng.dirichlet([a] * 4) for a in alpha])
        p = (p + 0.003) / 1.012
        w = np.round(np.log(p / 0.25) * 1000.0).astype(np.int32)
        ms = score_threshold(w, pvalue)
        pats.append({"weights": w, "min_score": ms, "pattern_id": i, "direction": DIR_P, "kind": PATTERN_PWM, "name": "SYN%04d" % i})
        if both_strands:  # reverse_complement, pattern.rs:103-112
            pats.append({"weights": np.ascontiguousarray(w[::-1, ::-1]), "min_score": ms, "pattern_id": i, "direction": DIR_N,
                         "kind": PATTERN_PWM, "name": "SYN%04d" % i})
    return pats


# ---- cohort ------------------------------------------------------------------------------------------------
def _concat_ranges(lo, hi):
    """Indices of the concatenation of [lo[i], hi[i]) plus the owner i of each index."""
    cnt = (hi - lo).astype(np.int64)
    tot = int(cnt.sum())
    owner = np.repeat(np.arange(len(lo)), cnt)
    first = np.cumsum(cnt) - cnt
    idx = np.arange(tot) - np.repeat(first, cnt) + np.repeat(lo, cnt)
    return idx, owner


def make_cohort(n_samples, n_regions, seed, lmax_pattern, region_len=(200, 2000), gap=(100, 1500), variant_rate=1.0 / 35,
                frac_ins=0.05, frac_del=0.05, indel_max=10, n_runs=0, lowercase_frac=0.0, two_beds=False, same_pos_frac=0.0,
                ld_blocks=0, bed_b_frac=0.55):
    """Random genome + regions + variants with a 1/k allele-count spectrum and uniform carriers.

    lmax_pattern: largest pattern length (the halo of main.rs:404-407 is lmax_pattern - 1 on both sides).
    n_runs: number of N runs put into the genome; lowercase_frac: fraction of the genome soft-masked;
    two_beds: add a second, partly overlapping / nested BED set (exercises merge + the asymmetric inner-region rule);
    same_pos_frac: fraction of variants duplicated at the same position with another ALT (overlap / truncation stress);
    ld_blocks: if > 0, carriers are drawn from that many founder haplotype classes (long shared haplotypes) instead of
    uniformly, which makes many haplotypes identical inside a region like real cohorts do.
    """
    rng = np.random.default_rng(seed)
    S, H = n_samples, 2 * n_samples
    halo = lmax_pattern - 1
    lens = rng.integers(region_len[0], region_len[1] + 1, size=n_regions)
    gaps = rng.integers(gap[0], gap[1] + 1, size=n_regions)
    starts = np.cumsum(gaps + np.concatenate([[0], lens[:-1] + 1])) + halo + 64
    ends = starts + lens
    glen = int(ends[-1] + halo + 256) if n_regions else 1024
    genome = LETTERS[rng.integers(0, 4, size=glen)].copy()
    for _ in range(n_runs):
        a = int(rng.integers(0, max(1, glen - 600)))
        genome[a:a + int(rng.integers(5, 400))] = ord("N")
    if lowercase_frac > 0:
        n_seg = max(1, int(glen * lowercase_frac / 300))
        for _ in range(n_seg):
            a = int(rng.integers(0, max(1, glen - 300)))
            seg = genome[a:a + 300]
            seg[:] = np.where(seg < 91, seg + 32, seg)

    bed_a = list(zip(starts.tolist(), ends.tolist()))
    peak_map = [bed_a]
    if two_beds:
        bed_b = []
        for s, e in bed_a:
            u = rng.random() * 0.55 / bed_b_frac  # bed_b_frac = share of the first set's regions that get a partner in the second
            if u < 0.25:
                bed_b.append((s + (e - s) // 3, e + int(rng.integers(0, 300))))      # overlaps the right edge
            elif u < 0.4:
                bed_b.append((s + 10, e - 10))                                        # strictly inside: never selected (Q1)
            elif u < 0.5:
                bed_b.append((s, e))                                                  # identical range in the other file
            elif u < 0.55:
                bed_b.append((max(halo, s - int(rng.integers(1, 200))), s))           # touches the left edge
        if bed_b and rng.random() < 0.9:
            bed_b.append(bed_b[0])                                                    # duplicate line: multiplicity 2 (Q3)
        peak_map.append(bed_b)
    merged = merge_regions([r for pm in peak_map for r in pm])
    m_start = np.array([m[0] for m in merged], dtype=np.int64)
    m_end = np.array([m[1] for m in merged], dtype=np.int64)
    R = len(merged)
    w_start = m_start - halo
    w_end = m_end + halo

    # variants
    n_var = int(rng.poisson(variant_rate * glen))
    pos = np.unique(rng.integers(1, glen - indel_max - 2, size=n_var)).astype(np.int64)
    if same_pos_frac > 0 and len(pos):
        extra = rng.choice(pos, size=int(len(pos) * same_pos_frac))
        pos = np.sort(np.concatenate([pos, extra]))
    n_var = len(pos)
    u = rng.random(n_var)
    kind = np.where(u < frac_ins, 1, np.where(u < frac_ins + frac_del, 2, 0))  # 0 SNV, 1 insertion, 2 deletion
    k_len = rng.integers(1, indel_max + 1, size=n_var)
    ref_len = np.where(kind == 2, k_len + 1, 1).astype(np.uint32)
    alt_len = np.where(kind == 1, k_len + 1, 1).astype(np.uint32)
    tot = (ref_len + alt_len).astype(np.int64)
    off = np.cumsum(tot) - tot
    allele = np.zeros(int(tot.sum()), dtype=np.uint8)
    ridx, rown = _concat_ranges(pos, pos + ref_len)
    allele[(off[rown] + (ridx - pos[rown]))] = genome[ridx]
    aoff = off + ref_len
    first = genome[pos]
    code = np.searchsorted(LETTERS, np.where(first > 90, first - 32, first))
    code = np.where((first == ord("N")) | (first == ord("n")), 0, code) % 4
    snv_alt = LETTERS[(code + rng.integers(1, 4, size=n_var)) % 4]
    allele[aoff] = np.where(kind == 0, snv_alt, first)
    iidx, iown = _concat_ranges(aoff + 1, aoff + alt_len)
    allele[iidx] = LETTERS[rng.integers(0, 4, size=len(iidx))]

    # carriers: allele count k with P(k) ~ 1/k, carriers uniform (drawn with replacement, duplicates collapse)
    ks = np.arange(1, max(2, H))
    pk = (1.0 / ks) / (1.0 / ks).sum()
    k = rng.choice(ks, size=n_var, p=pk)
    pitch = max(1, (H + 31) // 32)
    carriers = np.zeros((max(1, n_var), pitch), dtype=np.uint32)
    founders = None
    if ld_blocks > 0:
        founders = rng.integers(0, ld_blocks, size=H)
    chunk = max(1, (64 << 20) // max(1, pitch * 32))
    for c0 in range(0, n_var, chunk):
        c1 = min(n_var, c0 + chunk)
        kk = k[c0:c1]
        rows = np.repeat(np.arange(c1 - c0), kk)
        bits = np.zeros((c1 - c0, pitch * 32), dtype=bool)
        if founders is None:
            cols = rng.integers(0, H, size=len(rows))
            bits[rows, cols] = True
        else:
            # a variant is carried by whole founder classes: frequency ~ k / H of the classes
            nf = np.maximum(1, (kk * ld_blocks) // max(1, H))
            frows = np.repeat(np.arange(c1 - c0), nf)
            fcls = rng.integers(0, ld_blocks, size=len(frows))
            fm = np.zeros((c1 - c0, ld_blocks), dtype=bool)
            fm[frows, fcls] = True
            bits[:, :H] = fm[:, founders]
        carriers[c0:c1] = np.packbits(bits, axis=1, bitorder="little").view(np.uint32)

    # per region: records the BCF fetch of [w_start, w_end + 1) returns (haplotype.rs:79): overlap of [pos, pos + rlen)
    lo = np.searchsorted(pos, w_start - indel_max - 1, side="left")
    hi = np.searchsorted(pos, w_end, side="right")
    vidx, vown = _concat_ranges(lo, hi)
    keep = (pos[vidx] + ref_len[vidx] > w_start[vown]) & (pos[vidx] <= w_end[vown])
    vidx, vown = vidx[keep], vown[keep]
    var_off = np.zeros(R + 1, dtype=np.uint32)
    np.cumsum(np.bincount(vown, minlength=R), out=var_off[1:])
    variants = np.zeros(len(vidx), dtype=VARIANT_DTYPE)
    variants["pos"] = pos[vidx]
    variants["ref_off"] = off[vidx]
    variants["ref_len"] = ref_len[vidx]
    variants["alt_off"] = aoff[vidx]
    variants["alt_len"] = alt_len[vidx]
    variants["carrier_row"] = vidx

    # reference windows (FASTA fetch [start, end + 1), main.rs:157), clipped at the contig end
    w_hi = np.minimum(w_end + 1, glen)
    gidx, _ = _concat_ranges(w_start, w_hi)
    ref_bases = genome[gidx]
    ref_off = np.zeros(R + 1, dtype=np.uint64)
    np.cumsum(w_hi - w_start, out=ref_off[1:])

    inner_rows = []
    inner_off = np.zeros(R + 1, dtype=np.uint32)
    if len(peak_map) == 1 and len(merged) == len(bed_a):
        inner = np.zeros(R, dtype=INNER_DTYPE)
        inner["start"], inner["end"], inner["bed_index"], inner["multiplicity"] = m_start, m_end, 0, 1
        inner_off[:] = np.arange(R + 1)
    else:
        pm_arr = [(np.array([p[0] for p in pm], dtype=np.int64), np.array([p[1] for p in pm], dtype=np.int64)) for pm in peak_map]
        for r, m in enumerate(merged):
            near = []
            for b, (ps, pe) in enumerate(pm_arr):  # file order is kept: np.nonzero is ascending
                hit = ((ps <= m[0]) & (m[0] <= pe)) | ((ps <= m[1]) & (m[1] <= pe)) if len(ps) else np.zeros(0, dtype=bool)
                near.append([peak_map[b][i] for i in np.nonzero(hit)[0]])
            inner_rows.extend(select_inner_peaks(m, near))
            inner_off[r + 1] = len(inner_rows)
        inner = np.zeros(len(inner_rows), dtype=INNER_DTYPE)
        if inner_rows:
            arr = np.array(inner_rows, dtype=np.int64)
            inner["start"], inner["end"], inner["bed_index"], inner["multiplicity"] = arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3]

    blk = Block(S, w_start, w_end, ref_off, ref_bases, inner_off, inner, var_off, variants, allele, carriers)
    blk.meta = {"genome_len": glen, "n_variants_total": n_var, "merged": merged, "peak_map": peak_map, "halo": halo, "genome": genome,
                "var_pos": pos, "var_ref_off": off, "var_ref_len": ref_len, "var_alt_off": aoff, "var_alt_len": alt_len, "allele": allele}
    return blk


# BASELINE.json configs made concrete (sizes can be scaled down for tests)
def config2(scale=1.0, seed=2):
    """100 samples x 10k DHS regions (200-2000 bp) x 50 random PWMs (both strands), single GPU."""
    pats = make_pwms(50, seed=1000 + seed)
    lmax = max(p["weights"].shape[0] for p in pats)
    blk = make_cohort(100, max(1, int(10000 * scale)), seed=seed, lmax_pattern=lmax)
    return pats, blk


def config3(scale=1.0, seed=3, n_pwms=401):
    """BASELINE.json configs[2]: 2,504 samples (1000G-like) x 2 BED sets x 5,000 regions each (partly overlapping) x the HOCOMOCO
    v11 core collection size (401 PWMs, both strands; SURVEY D7: the count is a parameter).  The genome only spans the regions."""
    pats = make_pwms(n_pwms, seed=3000 + seed, lmin=7, lmax=25)
    lmax = max(p["weights"].shape[0] for p in pats)
    blk = make_cohort(2504, max(1, int(5000 * scale)), seed=seed, lmax_pattern=lmax, two_beds=True, n_runs=20, bed_b_frac=1.0)
    return pats, blk


def config4(n_regions=32, n_samples=100000, seed=4, n_pwms=50):
    """BASELINE.json configs[3]: biobank scale, 200,000 haplotypes, a record every ~4 bp (1/k allele-count spectrum, uniform
    carriers), 50 PWMs (L 8-30, both strands).  The full config has 100k regions over chr1; a bench step takes `n_regions` of them
    (regions are independent, main.rs:395-429) and runs them once per SAMPLE BLOCK (sharding.sample_block)."""
    pats = make_pwms(n_pwms, seed=4000 + seed)
    lmax = max(p["weights"].shape[0] for p in pats)
    blk = make_cohort(n_samples, n_regions, seed=seed, lmax_pattern=lmax, variant_rate=1.0 / 4, gap=(100, 300))
    return pats, blk
