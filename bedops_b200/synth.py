"""Seeded synthetic sorted BED of the shape BASELINE.json names (hg38 chromosome sizes, log-normal lengths).

Recipe = SURVEY.md 8(d): chromosomes chr1-22,X,Y emitted in strcmp order; per chromosome
n_c = round(N*size_c/sum(size)); start ~ U[0,size_c-1); len = max(1, floor(LogNormal(mu, sigma)));
end = min(start+len, size_c) forced > start; rows sorted by (start, end); columns chrom start end id<k> score,
score = integer U[0,1000).  numpy default_rng(seed), one generator per file.
"""
from __future__ import annotations

import numpy as np

HG38 = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}
MAP_SHAPE = (5.5, 1.0)   # median ~245 bp
REF_SHAPE = (7.0, 1.0)   # median ~1.1 kb


def columns(n_total: int, seed: int, mu: float, sigma: float, chroms=None, unique: bool = False):
    """Yield (chrom, start[int64], end[int64], score[int64]) per chromosome in strcmp order."""
    sizes = HG38 if chroms is None else {c: HG38[c] for c in chroms}
    total = float(sum(HG38.values()))
    rng = np.random.default_rng(seed)
    for c in sorted(sizes):
        m = int(round(n_total * sizes[c] / total))
        if m == 0:
            continue
        length = np.maximum(1, rng.lognormal(mu, sigma, m).astype(np.int64))
        s = rng.integers(0, sizes[c] - 1, m)
        e = np.minimum(s + length, sizes[c])
        e[e <= s] = s[e <= s] + 1
        order = np.lexsort((e, s))
        sc = rng.integers(0, 1000, m)
        s, e = s[order], e[order]
        if unique:
            keep = np.ones(m, dtype=bool)
            keep[1:] = (s[1:] != s[:-1]) | (e[1:] != e[:-1])
            s, e, sc = s[keep], e[keep], sc[: int(keep.sum())]
        yield c, s, e, sc


def bed_text(n_total: int, seed: int, shape=MAP_SHAPE, fields: int = 5, chroms=None, unique: bool = False,
             float_scores: bool = False) -> bytes:
    """BED3/BED5 text of ~n_total rows."""
    out = []
    k = 0
    frng = np.random.default_rng(seed + 1000)
    for c, s, e, sc in columns(n_total, seed, shape[0], shape[1], chroms, unique):
        m = len(s)
        if fields == 3:
            lines = ["%s\t%d\t%d" % (c, a, b) for a, b in zip(s.tolist(), e.tolist())]
        else:
            if float_scores:
                vals = frng.lognormal(0, 3, m) * frng.choice([-1.0, 1.0], m)
                scs = ["%.6f" % v for v in vals]
            else:
                scs = [str(v) for v in sc.tolist()]
            lines = ["%s\t%d\t%d\tid%d\t%s" % (c, a, b, k + i, v)
                     for i, (a, b, v) in enumerate(zip(s.tolist(), e.tolist(), scs))]
        k += m
        out.append("\n".join(lines))
    return ("\n".join(out) + "\n").encode() if out else b""
