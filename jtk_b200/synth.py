"""Deterministic synthetic inputs for the per-chunk pair-HMM path.

Shapes follow the reference's own synthetic drivers: `sandbox/src/bin/benchmark_clustering.rs:55-100`
(random template, haplotypes that differ by a few variants, reads with sub/ins/del at error/3 each) and
`sandbox/src/bin/gen_sim_genome.rs:23-28` (diploid mock genome).  Guide ops are the true generating
alignment (a valid global path: sum(non-Ins)=len(template), sum(non-Del)=len(read), the invariant of
`haplotyper/src/consensus/mod.rs:461-464`).  numpy only; no torch, no oracle.
"""
from __future__ import annotations

import numpy as np

OP_MATCH, OP_MISMATCH, OP_INS, OP_DEL = 0, 1, 2, 3
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_template(rng: np.random.Generator, length: int) -> np.ndarray:
    """Uniform random ACGT (ASCII uint8)."""
    return ACGT[rng.integers(0, 4, size=length)]


def mutate_read(rng: np.random.Generator, template: np.ndarray, error_rate: float):
    """One ONT-like read of `template`: per-base sub/ins/del each at error_rate/3.

    Returns (read ASCII uint8, ops uint8) where ops is the generating global alignment.
    """
    L = len(template)
    p = error_rate / 3.0
    u = rng.random(L)
    kind = np.where(u < p, OP_MISMATCH, np.where(u < 2 * p, OP_DEL, OP_MATCH)).astype(np.uint8)
    # number of inserted bases before each template position (and after the last): geometric
    n_ins = rng.geometric(1.0 - p, size=L + 1) - 1
    code = np.searchsorted(ACGT, template)
    sub = (code + rng.integers(1, 4, size=L)) % 4
    base = np.where(kind == OP_MISMATCH, ACGT[sub], template)
    # columns: for position j: n_ins[j] Ins columns then the template column
    tot = int(n_ins.sum()) + L
    ops = np.empty(tot, dtype=np.uint8)
    bases = np.empty(tot, dtype=np.uint8)
    col_start = np.cumsum(n_ins[:L] + 1) - 1  # index of the template column of position j
    ops[:] = OP_INS
    bases[:] = ACGT[rng.integers(0, 4, size=tot)]
    ops[col_start] = kind
    bases[col_start] = base
    keep = ops != OP_DEL
    read = bases[keep]
    if len(read) == 0:  # degenerate; never for realistic sizes
        return template.copy(), np.zeros(L, dtype=np.uint8)
    return read.copy(), ops


def diploid_chunk(seed: int, length: int = 2000, n_reads: int = 60, error_rate: float = 0.08,
                  n_snv: int = 5, n_hap: int = 2):
    """BASELINE.json configs[0]: one chunk, n_reads reads from n_hap haplotypes differing by SNVs.

    Returns dict(template, reads[list], ops[list], strands uint8[n], hap int[n], snv_pos).
    SNVs are >= 7 bp from the ends and from each other (MASK_LENGTH, pseudo_mcmc.rs:3,445,548)
    and avoid homopolymers longer than 2 (pseudo_mcmc.rs:4,509-511).
    """
    rng = np.random.default_rng(seed)
    tmpl = random_template(rng, length)
    haps = [tmpl]
    snvs = []
    for _ in range(1, n_hap):
        h = haps[0].copy()
        pos = []
        tries = 0
        while len(pos) < n_snv and tries < 10000:
            tries += 1
            j = int(rng.integers(20, length - 20))
            if any(abs(j - x) < 20 for x in pos + [p for s in snvs for p in s]):
                continue
            if tmpl[j] == tmpl[j - 1] or tmpl[j] == tmpl[j + 1]:
                continue
            c = int(np.searchsorted(ACGT, tmpl[j]))
            nb = ACGT[(c + int(rng.integers(1, 4))) % 4]
            if nb == tmpl[j - 1] or nb == tmpl[j + 1]:
                continue
            h[j] = nb
            pos.append(j)
        haps.append(h)
        snvs.append(sorted(pos))
    reads, ops, hap_of = [], [], []
    per = n_reads // n_hap
    for hi, h in enumerate(haps):
        cnt = per if hi < n_hap - 1 else n_reads - per * (n_hap - 1)
        for _ in range(cnt):
            r, o = mutate_read(rng, h, error_rate)
            # ops were generated against haplotype h; they are also a valid global path against
            # `tmpl` because the haplotypes differ by substitutions only.
            reads.append(r)
            ops.append(o)
            hap_of.append(hi)
    strands = (rng.random(len(reads)) < 0.5).astype(np.uint8)
    return dict(template=tmpl, reads=reads, ops=ops, strands=strands,
                hap=np.array(hap_of, dtype=np.int32), snv_pos=snvs, haps=haps)


def paralog_chunk(seed: int, length: int = 2000, n_reads: int = 240, error_rate: float = 0.08,
                  n_paralog: int = 4, paralog_div: float = 0.05, hap_div: float = 0.001):
    """BASELINE.json configs[3]: a repeat-heavy chunk -- n_paralog copies that differ from the first by `paralog_div`
    substitutions per base (`sandbox/src/bin/gen_sim_genome_segdup.rs:30-35` style), each with two haplotypes at `hap_div`;
    n_reads reads split evenly over the 2 * n_paralog sequences.  Substitutions only, so the generating alignment of a read
    is a valid global path against `template` (copy 0, haplotype 0).
    Returns dict(template, reads, ops, strands, paralog int[n], hap int[n])."""
    rng = np.random.default_rng(seed)
    tmpl = random_template(rng, length)

    def substitute(seq, rate, lo=10):
        out = seq.copy()
        n = max(1, int(round(rate * length)))
        pos = rng.choice(np.arange(lo, length - lo), size=n, replace=False)
        out[pos] = ACGT[(np.searchsorted(ACGT, out[pos]) + rng.integers(1, 4, size=n)) % 4]
        return out
    seqs = []
    for p in range(n_paralog):
        a = tmpl if p == 0 else substitute(tmpl, paralog_div)
        seqs.append((p, a))
        seqs.append((p, substitute(a, hap_div)))
    reads, ops, par, hap = [], [], [], []
    per = n_reads // len(seqs)
    for h, (p, sq) in enumerate(seqs):
        cnt = per if h < len(seqs) - 1 else n_reads - per * (len(seqs) - 1)
        for _ in range(cnt):
            r, o = mutate_read(rng, sq, error_rate)
            reads.append(r); ops.append(o); par.append(p); hap.append(h)
    strands = (rng.random(len(reads)) < 0.5).astype(np.uint8)
    return dict(template=tmpl, reads=reads, ops=ops, strands=strands, paralog=np.array(par, dtype=np.int32),
                hap=np.array(hap, dtype=np.int32))


def diploid_region(seed: int, n_chunks: int, length: int = 2000, n_reads: int = 60,
                   error_rate: float = 0.08, div: float = 0.0005):
    """BASELINE.json configs[1]/[2]: chunks of a mock diploid region (hap B = hap A + `div` SNV rate,
    `sandbox/src/bin/gen_sim_genome.rs:23-28`), n_reads per chunk split over the two haplotypes."""
    out = []
    for c in range(n_chunks):
        n_snv = max(1, int(round(div * length))) if div > 0 else 0
        out.append(diploid_chunk(seed * 1000003 + c, length, n_reads, error_rate, n_snv=n_snv))
    return out


def cell_count(ops: np.ndarray, Lt: int, Lr: int, radius: int) -> int:
    """C = sum_d w(d) for one pair -- numpy restatement of the band geometry."""
    nd = Lt + Lr + 1
    centre = np.zeros(nd, dtype=np.int64)
    is_m = ops <= OP_MISMATCH
    di = np.where(ops == OP_DEL, 0, 1)
    step = np.where(is_m, 2, 1)
    d_after = np.cumsum(step)
    i_after = np.cumsum(di)
    centre[d_after] = i_after
    # skipped diagonals of match steps keep the previous row
    dm = d_after[is_m] - 1
    centre[dm] = i_after[is_m] - 1
    d = np.arange(nd)
    lo = np.maximum(np.maximum(centre - radius, 0), d - Lt)
    hi = np.minimum(np.minimum(centre + radius, Lr), d)
    return int(np.maximum(hi - lo + 1, 0).sum())
