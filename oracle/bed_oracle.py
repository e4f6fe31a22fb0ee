"""bed_oracle.py -- CPU restatement of the BEDOPS v2.4.26 sorted-interval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under bedops_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg use it, and only as the checker.

Parity status: PINNED.  tests/test_oracle.py checks every function here against
  * the reference's own golden vectors (applications/bed/bedops/test/TestPlan.xml, orders listed in
    SURVEY.md 8c) and the worked examples in docs/content/reference/statistics/bedmap.rst, committed as
    fixtures under tests/golden/ by tests/golden/make_golden.py, and
  * outputs of the unmodified reference binaries built by oracle/build_ref.sh (oracle/_ref/bin), on seeded
    synthetic inputs (fixtures committed; live differential runs when the binaries are present).

Each function cites the reference code it restates (paths relative to the reference root).
"""
from __future__ import annotations

import bisect
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

WS = b" \t\r\v\f"


@dataclass
class Row:
    chrom: bytes
    start: int
    end: int
    rest: bytes          # verbatim remainder of the line after the last parsed field (leading tab included)
    id: bytes = b""
    score: float = 0.0
    rest3: bytes = b""   # remainder after column 3 (what B3Rest keeps)


def _token(line: bytes, pos: int) -> Tuple[bytes, int]:
    n = len(line)
    while pos < n and line[pos:pos + 1] in (b" ", b"\t", b"\r", b"\v", b"\f"):
        pos += 1
    s = pos
    while pos < n and line[pos:pos + 1] not in (b" ", b"\t", b"\r", b"\v", b"\f"):
        pos += 1
    return line[s:pos], pos


def _number(line: bytes, pos: int) -> Tuple[bytes, int]:
    """Longest prefix strtod would accept (decimal form)."""
    n = len(line)
    s = pos
    if pos < n and line[pos:pos + 1] in (b"+", b"-"):
        pos += 1
    while pos < n and line[pos:pos + 1].isdigit():
        pos += 1
    if pos < n and line[pos:pos + 1] == b".":
        pos += 1
        while pos < n and line[pos:pos + 1].isdigit():
            pos += 1
    if pos < n and line[pos:pos + 1] in (b"e", b"E"):
        q = pos + 1
        if q < n and line[q:q + 1] in (b"+", b"-"):
            q += 1
        if q < n and line[q:q + 1].isdigit():
            while q < n and line[q:q + 1].isdigit():
                q += 1
            pos = q
    return line[s:pos], pos


def parse_bed(text: bytes, min_fields: int = 3) -> List[Row]:
    """Bed::B3Rest/B4Rest/B5Rest::readline(FILE*)  (interfaces/general-headers/data/bed/Bed.hpp:343-360,
    :577-606, :829-860) with formats "%s\\t%lu\\t%lu%[^\\n]s\\n" etc. (:380-382, :644-646, :901-903).
    A last line without '\\n' is not a record (allocate_iterator_starch_bed tests feof after the read,
    data/bed/AllocateIterator_BED_starch.hpp:172-187); blank lines are skipped by %s."""
    rows: List[Row] = []
    end = text.rfind(b"\n")
    if end < 0:
        return rows
    for line in text[:end].split(b"\n"):
        if not line.strip(WS):
            continue
        chrom, p = _token(line, 0)
        st, p = _token(line, p)
        en, p = _token(line, p)
        # %lu stops at the first non-digit: the rest starts right there
        sdig = st.lstrip(b"+")
        edig_full = en.lstrip(b"+")
        k = 0
        while k < len(edig_full) and edig_full[k:k + 1].isdigit():
            k += 1
        p -= len(edig_full) - k
        r = Row(chrom, int(sdig), int(edig_full[:k]), line[p:], rest3=line[p:])
        if min_fields >= 4:
            r.id, p = _token(line, p)
            r.rest = line[p:]
            if min_fields >= 5:
                while p < len(line) and line[p:p + 1] in (b" ", b"\t", b"\r", b"\v", b"\f"):
                    p += 1
                num, p = _number(line, p)
                r.score = float(num)
                r.rest = line[p:]
        rows.append(r)
    return rows


def by_chrom(rows: Sequence[Row]) -> Dict[bytes, List[Row]]:
    out: Dict[bytes, List[Row]] = {}
    for r in rows:
        out.setdefault(r.chrom, []).append(r)
    return out


def chrom_order(*files: Sequence[Row]) -> List[bytes]:
    """strcmp order of chromosome names (BedCheckIterator.hpp:596-597)."""
    names = set()
    for f in files:
        for r in f:
            names.add(r.chrom)
    return sorted(names)


# --------------------------------------------------------------------------------------------------
# bedmap
# --------------------------------------------------------------------------------------------------
DBL_EPS = 2.220446049250313e-16


def _perc(p: float) -> float:
    """PercentOverlapMapping constructor (data/bed/BedDistances.hpp:120-131)."""
    while p > 1:
        p /= 10.0
    p -= DBL_EPS
    if p <= 0.0:
        p = DBL_EPS
    return p


def qualifies(kind: str, val, rs: int, re: int, ms: int, me: int) -> Tuple[bool, int]:
    """dist_.Map2Ref(m, r) == 0 for the overlap criteria of Bedmap.cpp:107-156 / BedDistances.hpp:41-317.
    Returns (qualifies, overlap bases)."""
    ov = max(0, min(re, me) - max(rs, ms))
    if kind == "bp":        # Bed::Overlapping(N), :96-115
        return ov > 0 and ov >= val, ov
    if kind == "range":     # Bed::RangedDist(N), :41-67
        if ms < re:
            return me + val > rs, ov
        return re + val > ms, ov
    if kind == "exact":     # Bed::Exact, :285-308
        return rs == ms and re == me, ov
    if ov <= 0:
        return False, ov
    p = _perc(val)
    fm = ov / float(me - ms) >= p   # PercentOverlapMapping::Ref2Map, :136-172
    fr = ov / float(re - rs) >= p   # PercentOverlapReference, :190-213
    if kind == "fraction-map":
        return fm, ov
    if kind == "fraction-ref":
        return fr, ov
    if kind == "fraction-either":
        return fm or fr, ov
    if kind == "fraction-both":
        return fm and fr, ov
    raise ValueError(kind)


def _fmt_score(v: float, prec: int, sci: bool) -> bytes:
    """Formats::Format(double, precision, scientific) -> "%.<prec>lf" / "%.<prec>e" (utility/Formats.hpp:42-49)."""
    return (("%%.%de" if sci else "%%.%df") % prec % v).encode()


def echo_row(r: Row, fields: int) -> bytes:
    """B4Rest/B5Rest::print (Bed.hpp:640-646 "%s\\t%lu\\t%lu\\t%s%s", :740-743/:896-903 "...\\t%s\\t%lf%s"): used when the
    reference file has the map's record type, i.e. single-file bedmap (Bedmap.cpp:676-700)."""
    if fields <= 3:
        return echo_b3rest(r)
    out = r.chrom + b"\t" + str(r.start).encode() + b"\t" + str(r.end).encode() + b"\t" + r.id
    if fields >= 5:
        out += b"\t" + (b"%f" % r.score)
    return out + r.rest


def echo_b3rest(r: Row) -> bytes:
    """B3Rest::print "%s\\t%lu\\t%lu%s" (Bed.hpp:316-320, :376-378)."""
    return r.chrom + b"\t" + str(r.start).encode() + b"\t" + str(r.end).encode() + r.rest3


class NanElement(Exception):
    """--max-element / --min-element on a reference row without mapped elements: the reference throws "Unable to process
    a 'NAN' with PrintAllScorePrecision." (ProcessBedVisitorRow.hpp:206-208); .stdout is what it had printed by then."""

    def __init__(self, stdout: bytes):
        super().__init__("Unable to process a 'NAN' with PrintAllScorePrecision.")
        self.stdout = stdout


def bedmap(ref_text: bytes, map_text: Optional[bytes], ops: Sequence[str], overlap: Tuple[str, object] = ("bp", 1),
           prec: int = 6, sci: bool = False, delim: bytes = b"|", multidelim: bytes = b";",
           skip_unmapped: bool = False, chrom: Optional[bytes] = None) -> bytes:
    """WindowSweep::sweep + BedBaseVisitor + MultiVisitor (interfaces/src/algorithm/sweep/WindowSweepImpl.cpp:174-256,
    algorithm/visitors/bed/BedBaseVisitor.hpp:118-225, visitors/other/MultiVisitor.hpp:83-98), restated
    declaratively: for every reference row the qualifying map rows are {m : dist_.Map2Ref(m, r) == 0}, visited in
    file order.  Sums are taken fresh per reference row (the reference keeps a rolling double, SumVisitor.hpp:47-51;
    identical for exactly representable partial sums, see DESIGN.md parity notes)."""
    need_fields = 3
    for o in ops:
        if o in ("sum", "mean", "max", "min", "echo-map-score", "variance", "stdev", "cv", "median", "wmean", "max-element",
                 "min-element") or o.startswith("kth:") or o.startswith("mad") or o.startswith("tmean:"):
            need_fields = max(need_fields, 5)   # Input.hpp:404-420: the map record type is the widest any visitor needs
        elif o in ("echo-map-id", "echo-map-id-uniq"):
            need_fields = max(need_fields, 4)
    single = map_text is None
    maps = parse_bed(ref_text if single else map_text, need_fields)
    refs = maps if single else parse_bed(ref_text, 3)
    mchrom = by_chrom(maps)
    kind, val = overlap
    pad = val if kind == "range" else 0
    out: List[bytes] = []
    cache = {}
    rowid = 0
    for ri, r in enumerate(refs):
        if chrom is not None and chrom != b"all" and r.chrom != chrom:
            continue
        if r.chrom not in cache:
            ml = mchrom.get(r.chrom, [])
            starts = [m.start for m in ml]
            pm, cur = [], 0
            for m in ml:
                cur = max(cur, m.end)
                pm.append(cur)
            cache[r.chrom] = (ml, starts, pm)
        ml, starts, pm = cache[r.chrom]
        hi = bisect.bisect_left(starts, r.end + pad)
        lo = bisect.bisect_left(pm, r.start - pad + 1, 0, hi) if r.start >= pad else 0
        hits = []
        bases = 0
        for k in range(lo, hi):
            m = ml[k]
            ok, ov = qualifies(kind, val, r.start, r.end, m.start, m.end)
            if ok:
                hits.append(m)
                bases += ov
        cnt = len(hits)
        if skip_unmapped and cnt == 0:   # MultiVisitor.hpp:84-85
            continue
        cols: List[bytes] = []
        for o in ops:
            if o == "echo":              # EchoVisitor.hpp:39-65
                cols.append(echo_row(r, need_fields if single else 3))
            elif o == "count":           # CountVisitor.hpp:34-64
                cols.append(str(cnt).encode())
            elif o == "indicator":       # IndicatorVisitor.hpp:37-56
                cols.append(b"1" if cnt else b"0")
            elif o == "bases":           # OvrAggregateVisitor.hpp:41-108
                cols.append(str(bases).encode())
            elif o in ("sum", "mean", "max", "min"):
                if cnt == 0:
                    cols.append(b"NAN")  # Signal::NaN (interfaces/src/data/measurement/NaN.cpp:27)
                else:
                    if o == "sum":       # SumVisitor.hpp:36-68
                        s = 0.0
                        for m in hits:
                            s += m.score
                        v = s
                    elif o == "mean":    # AverageVisitor.hpp:35-76: sum_/counter_
                        s = 0.0
                        for m in hits:
                            s += m.score
                        v = s / cnt
                    elif o == "max":     # ExtremeVisitor.hpp:84-134
                        v = max(m.score for m in hits)
                    else:
                        v = min(m.score for m in hits)
                    cols.append(_fmt_score(v, prec, sci))
            elif o == "echo-map-id":     # EchoMapBedVisitor.hpp:39-66 (ties in file order, SURVEY hazard 2)
                cols.append(multidelim.join(m.id for m in hits))
            elif o == "median" or o.startswith("kth:"):
                # RollingKthAverage over RollingKth (RollingKthAverageVisitor.hpp:37-68, RollingKthVisitor.hpp:75-95):
                # the marker sits at iround(k*n)-1; a "true integer" k*n averages two neighbours -- which also happens
                # whenever floor(k*n) == 0, because positions are only decremented when they are positive
                k = 0.5 if o == "median" else float(o[4:])
                v = sorted(m.score for m in hits)
                n = len(v)
                if n == 0:
                    cols.append(b"NAN")
                elif n == 1:
                    cols.append(_fmt_score(v[0], prec, sci))
                else:
                    kn = k * n
                    up, down = math.ceil(kn), math.floor(kn)
                    d1 = math.ceil(kn)
                    pos = math.floor(kn) if d1 - kn > 0.5 else d1       # iround, RollingKthVisitor.hpp:117-124
                    up, down, pos = (x - 1 if x > 0 else x for x in (up, down, pos))
                    if up == down:
                        cols.append(_fmt_score((v[pos] + v[pos + 1]) / 2.0, prec, sci))
                    elif pos == up:
                        cols.append(_fmt_score(v[pos], prec, sci))
                    else:
                        cols.append(_fmt_score(v[pos + 1], prec, sci))
            elif o.startswith("mad"):
                # MedianAbsoluteDeviationVisitor.hpp:57-93: NAN for fewer than two hits; the median (as --median gives
                # it), then the plain median of |x - median| (mean of the two middle values for an even count) * mult
                mult = float(o[4:]) if o.startswith("mad:") else 1.0
                v = sorted(m.score for m in hits)
                n = len(v)
                if n <= 1:
                    cols.append(b"NAN")
                else:
                    med = (v[n // 2 - 1] + v[n // 2]) / 2.0 if n % 2 == 0 else v[(n - 1) // 2]
                    dv = sorted(abs(x - med) for x in v)
                    mad = (dv[n // 2 - 1] + dv[n // 2]) / 2.0 if n % 2 == 0 else dv[n // 2]
                    cols.append(_fmt_score(mad * mult, prec, sci))
            elif o in ("variance", "stdev", "cv"):  # VarianceVisitor.hpp:58-66, StdevVisitor.hpp, CoeffVariationVisitor.hpp
                sm = sq = 0.0
                for m in hits:
                    sm += m.score
                    sq += m.score * m.score
                n = float(cnt)
                if cnt <= 1:
                    cols.append(b"NAN")
                else:
                    var = ((n * sq) - (sm * sm)) / (n * (n - 1.0))
                    if o == "variance":
                        cols.append(_fmt_score(var, prec, sci))
                    elif o == "stdev":
                        cols.append(_fmt_score(math.sqrt(var), prec, sci))
                    elif sm / n == 0:
                        cols.append(b"NAN")
                    else:
                        cols.append(_fmt_score(math.sqrt(var) / (sm / n), prec, sci))
            elif o == "wmean":
                # WeightedAverageVisitor.hpp:55-70: sum(w*score)/sum(w), w = overlap/length of the reference row.  The
                # reference adds in heap-address order; file order here (equal up to the rounding of the two sums)
                if cnt == 0:
                    cols.append(b"NAN")
                else:
                    value = wsum = 0.0
                    for m in hits:
                        w = max(0, min(r.end, m.end) - max(r.start, m.start)) / float(r.end - r.start)
                        value += w * m.score
                        wsum += w
                    cols.append(b"-nan" if wsum == 0.0 and value == 0.0 else _fmt_score(value / wsum, prec, sci))
            elif o.startswith("tmean:"):
                # TrimmedMeanVisitor.hpp:93-141 with both markers freshly positioned (the reference leaves the lower
                # marker where earlier rows put it when round(low*n) == 0: see DESIGN.md parity notes)
                low, high = (float(x) for x in o[6:].split(":"))
                v = sorted(m.score for m in hits)
                n = len(v)
                if n == 0:
                    cols.append(b"NAN")
                else:
                    def iround(d):
                        d1 = math.ceil(d)
                        return math.floor(d) if d1 - d > 0.5 else d1
                    do_kth = abs(1.0 - low - high) <= DBL_EPS
                    pl, ph = int(iround(low * n)), n - int(iround(high * n))
                    if abs(low - high) <= DBL_EPS:
                        pl = max(pl, n - ph)
                        ph = n - pl
                    do_low = pl > 0
                    if do_low:
                        pl -= 1
                    if ph > 0:
                        ph -= 1
                    if do_kth or (do_low and ph == pl):
                        cols.append(_fmt_score(v[ph], prec, sci))
                    else:
                        # file order, like the device (the reference's rolling sums are exact for integer scores)
                        ranked = sorted(range(n), key=lambda k: (hits[k].score, k))
                        keep = set(ranked[pl + 1:ph + 1] if do_low else ranked[:ph + 1])
                        sm = 0.0
                        for k in range(n):
                            if k in keep:
                                sm += hits[k].score
                        cols.append(_fmt_score(sm / (ph - pl) if do_low else sm / (ph + 1), prec, sci))
            elif o in ("max-element", "min-element"):
                # ExtremeVisitor.hpp:84-134 ordered by ScoreThenGenomicCompare{Greater,Lesser} (BedCompare.hpp:263-288): best
                # score, then the genomically last (max) / first (min) row, then the first in file order; printed by
                # PrintAllScorePrecision (ProcessBedVisitorRow.hpp:181-222), which throws on an empty set
                if cnt == 0:
                    out.append(delim.join(cols) + (delim if cols else b""))
                    raise NanElement(b"".join(out))
                best = hits[0]
                for m in hits[1:]:
                    if o == "max-element":
                        better = (m.score, m.start, m.end) > (best.score, best.start, best.end)
                    else:
                        better = (m.score, m.start, m.end) < (best.score, best.start, best.end)
                    if better:
                        best = m
                cols.append(best.chrom + b"\t%d\t%d\t" % (best.start, best.end) + best.id + b"\t" +
                            _fmt_score(best.score, prec, sci) + best.rest)
            elif o == "echo-map":        # EchoMapBed<PrintRangeDelim<PrintAll>>: the map rows as their record type prints them
                cols.append(multidelim.join(echo_row(m, need_fields) for m in hits))
            elif o == "echo-map-score":  # PrintRangeDelim<PrintScorePrecision> (ProcessBedVisitorRow.hpp:152-176)
                cols.append(multidelim.join(_fmt_score(m.score, prec, sci) for m in hits))
            elif o == "echo-map-size":   # PrintRangeDelim<PrintLength> (:309-316)
                cols.append(multidelim.join(str(m.end - m.start).encode() for m in hits))
            elif o == "echo-overlap-size":  # EchoMapIntersectLengthVisitor.hpp:64-73: length of ref.intersection(map), 0 if disjoint
                cols.append(multidelim.join(str(max(0, min(r.end, m.end) - max(r.start, m.start))).encode() for m in hits))
            elif o == "echo-map-range":  # PrintGenomicRange<PrintBED3> (:433-456): one range spanning every hit
                if hits:
                    cols.append(r.chrom + b"\t%d\t%d" % (min(m.start for m in hits), max(m.end for m in hits)))
                else:
                    cols.append(b"")
            elif o in ("bases-uniq", "bases-uniq-f"):  # OvrUniqueVisitor.hpp:62-77, OvrUniqueFractionVisitor.hpp:47-50
                ovr = 0
                if hits:
                    cs, ce = hits[0].start, hits[0].end
                    for m in hits[1:]:
                        if min(ce, m.end) > max(cs, m.start):       # Bed.hpp:172-191 overlap() > 0
                            cs, ce = min(cs, m.start), max(ce, m.end)
                        else:
                            ovr += max(0, min(ce, r.end) - max(cs, r.start))
                            cs, ce = m.start, m.end
                    ovr += max(0, min(ce, r.end) - max(cs, r.start))
                if o == "bases-uniq":
                    cols.append(str(ovr).encode())
                else:
                    cols.append(_fmt_score(float(ovr) / float(r.end - r.start), prec, sci))
            elif o == "echo-map-id-uniq":  # PrintUniqueRangeIDs (ProcessBedVisitorRow.hpp:361-389): std::set<std::string>
                cols.append(multidelim.join(sorted(set(m.id for m in hits))))
            elif o == "echo-ref-size":
                cols.append(str(r.end - r.start).encode())
            elif o == "echo-ref-name":
                cols.append(r.chrom + b":" + str(r.start).encode() + b"-" + str(r.end).encode())
            elif o == "echo-ref-row-id":
                rowid += 1   # PrintRowID: a static counter bumped by every printed id (ProcessBedVisitorRow.hpp:347-354)
                cols.append(b"id-" + str(rowid).encode())
            else:
                raise ValueError(o)
        out.append(delim.join(cols) + b"\n")
    return b"".join(out)


# --------------------------------------------------------------------------------------------------
# bedops
# --------------------------------------------------------------------------------------------------
def _sel(rows: Sequence[Row], chrom: Optional[bytes]) -> List[Row]:
    if chrom is None or chrom == b"all":
        return list(rows)
    return [r for r in rows if r.chrom == chrom]


def merged_union(files: Sequence[Sequence[Row]]) -> Dict[bytes, List[Tuple[int, int]]]:
    """nextMergeAllLines (applications/bed/bedops/src/Bedops.cpp:1186-1243): take the minimum (chrom,start) over
    all files, extend its end while any file's next element has start <= current end (touching coalesces, :1233)."""
    out: Dict[bytes, List[Tuple[int, int]]] = {}
    per: Dict[bytes, List[Tuple[int, int]]] = {}
    for f in files:
        for r in f:
            per.setdefault(r.chrom, []).append((r.start, r.end))
    for c, iv in per.items():
        iv.sort(key=lambda t: t[0])
        res: List[Tuple[int, int]] = []
        cs, ce = iv[0]
        for s, e in iv[1:]:
            if s <= ce:
                if e > ce:
                    ce = e
            else:
                res.append((cs, ce))
                cs, ce = s, e
        res.append((cs, ce))
        out[c] = res
    return out


def bedops_merge(texts: Sequence[bytes], chrom: Optional[bytes] = None) -> bytes:
    """doMerge (Bedops.cpp:592-606) -> record() -> B3NoRest::println "%s\\t%lu\\t%lu\\n" (Bed.hpp:228-232)."""
    files = [_sel(parse_bed(t, 3), chrom) for t in texts]
    u = merged_union(files)
    out = []
    for c in sorted(u):
        for s, e in u[c]:
            out.append(c + b"\t%d\t%d\n" % (s, e))
    return b"".join(out)


def bedops_intersect(texts: Sequence[bytes], chrom: Optional[bytes] = None) -> bytes:
    """doIntersection / nextIntersectLine (Bedops.cpp:574-587, :1105-1181): every file is merged within itself
    (getNextFileMergedCoords :791-814, mergeOverlap :864-886), then the common coverage of all files is emitted,
    positive length only."""
    lists = [merged_union([_sel(parse_bed(t, 3), chrom)]) for t in texts]
    out = []
    common = set(lists[0])
    for l in lists[1:]:
        common &= set(l)
    for c in sorted(common):
        acc = lists[0][c]
        for l in lists[1:]:
            b = l[c]
            res = []
            i = j = 0
            while i < len(acc) and j < len(b):
                lo = max(acc[i][0], b[j][0])
                hi = min(acc[i][1], b[j][1])
                if hi > lo:
                    res.append((lo, hi))
                if acc[i][1] < b[j][1]:
                    i += 1
                else:
                    j += 1
            acc = res
        for s, e in acc:
            out.append(c + b"\t%d\t%d\n" % (s, e))
    return b"".join(out)


def _subtract(a: List[Tuple[int, int]], b: List[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """pieces of the disjoint sorted list a that no segment of the disjoint sorted list b covers"""
    res, j = [], 0
    for s, e in a:
        while j < len(b) and b[j][1] <= s:
            j += 1
        cur, k = s, j
        while k < len(b) and b[k][0] < e:
            if b[k][0] > cur:
                res.append((cur, b[k][0]))
            cur = max(cur, b[k][1])
            k += 1
        if cur < e:
            res.append((cur, e))
    return res


def bedops_complement(texts: Sequence[bytes], full_left: bool = False, chrom: Optional[bytes] = None) -> bytes:
    """doComplement / nextComplementLine (Bedops.cpp:475-488, :891-943): the gaps between consecutive segments of the
    merged union of all files, per chromosome; with -L also [0, first start) when the first segment does not start at 0."""
    u = merged_union([_sel(parse_bed(t, 3), chrom) for t in texts])
    out = []
    for c in sorted(u):
        segs = u[c]
        if full_left and segs[0][0] != 0:
            out.append(c + b"\t0\t%d\n" % segs[0][0])
        for (s0, e0), (s1, e1) in zip(segs, segs[1:]):
            out.append(c + b"\t%d\t%d\n" % (e0, s1))
    return b"".join(out)


def bedops_difference(texts: Sequence[bytes], chrom: Optional[bytes] = None) -> bytes:
    """doDifference / nextDifferenceLine (Bedops.cpp:501-525, :948-1018): the first file merged within itself
    (getNextFileMergedCoords) minus the merged union of the other files; pieces of positive length."""
    a = merged_union([_sel(parse_bed(texts[0], 3), chrom)])
    b = merged_union([_sel(parse_bed(t, 3), chrom) for t in texts[1:]])
    out = []
    for c in sorted(a):
        for s, e in _subtract(a[c], b.get(c, [])):
            out.append(c + b"\t%d\t%d\n" % (s, e))
    return b"".join(out)


def bedops_symmdiff(texts: Sequence[bytes], chrom: Optional[bytes] = None) -> bytes:
    """doSymmetricDifference / nextSymmetricDiffLine (Bedops.cpp:698-745, :1341-1463): the bases covered by exactly one
    file (every file merged within itself), abutting pieces coalesced (mergeOverlap :864-886) =
    union of all files minus the union of the pairwise intersections."""
    lists = [merged_union([_sel(parse_bed(t, 3), chrom)]) for t in texts]
    chroms = sorted(set().union(*[set(l) for l in lists]))
    out = []
    for c in chroms:
        per = [l.get(c, []) for l in lists]
        cover = sorted(x for l in per for x in l)
        u: List[Tuple[int, int]] = []
        for s, e in cover:                      # union, touching coalesced
            if u and s <= u[-1][1]:
                u[-1] = (u[-1][0], max(u[-1][1], e))
            else:
                u.append((s, e))
        twice: List[Tuple[int, int]] = []
        for i in range(len(per)):
            for j in range(i + 1, len(per)):
                x, y, a, b = 0, 0, per[i], per[j]
                while x < len(a) and y < len(b):
                    lo, hi = max(a[x][0], b[y][0]), min(a[x][1], b[y][1])
                    if hi > lo:
                        twice.append((lo, hi))
                    if a[x][1] < b[y][1]:
                        x += 1
                    else:
                        y += 1
        twice.sort()
        t2: List[Tuple[int, int]] = []
        for s, e in twice:
            if t2 and s <= t2[-1][1]:
                t2[-1] = (t2[-1][0], max(t2[-1][1], e))
            else:
                t2.append((s, e))
        for s, e in _subtract(u, t2):
            out.append(c + b"\t%d\t%d\n" % (s, e))
    return b"".join(out)


def bedops_chop(texts: Sequence[bytes], chunk: int = 1, stagger: int = 0, exclude_short: bool = False,
                chrom: Optional[bytes] = None) -> bytes:
    """doChop (Bedops.cpp:438-467): every segment of the merged union is cut into pieces of `chunk` bases starting every
    `stagger` bases (every `chunk` bases when stagger is 0); a piece that would pass the segment end is clipped to it,
    or, with -x, ends the segment's pieces."""
    u = merged_union([_sel(parse_bed(t, 3), chrom) for t in texts])
    step = stagger if stagger else chunk
    out = []
    for c in sorted(u):
        for s, e in u[c]:
            i = s
            while i < e:
                pe = i + chunk
                if pe > e:
                    if exclude_short:
                        break
                    pe = e
                out.append(c + b"\t%d\t%d\n" % (i, pe))
                i += step
    return b"".join(out)


def bedops_everything(texts: Sequence[bytes], chrom: Optional[bytes] = None) -> bytes:
    """doUnionAll / nextUnionAllLine (Bedops.cpp:761-786, :1468-1516): multiset union of the rows of all files, every row
    printed as B3Rest (chrom, start, end re-printed, rest verbatim).  Order: chromosome, start, end, then strcmp of the
    rest of the line; on a full tie the row of the lower-numbered file comes first (the comparisons are strict)."""
    rows = []
    for f, t in enumerate(texts):
        for k, r in enumerate(_sel(parse_bed(t, 3), chrom)):
            rows.append((r.chrom, r.start, r.end, r.rest3, f, k, r))
    rows.sort(key=lambda x: x[:6])
    return b"".join(echo_b3rest(x[6]) + b"\n" for x in rows)


def bedops_element_of(texts: Sequence[bytes], thr: float = 1.0, use_pct: bool = True, invert: bool = False,
                      chrom: Optional[bytes] = None) -> bytes:
    """doElementOf / nextElementOfLine (Bedops.cpp:538-566, :1023-1100): overlap bases of each reference row
    (file 0, unmerged, all columns) with the union-merge of the other files; keep when
    rangeOverlap/range >= threshold (percent) or rangeOverlap >= threshold (bp) (:1094-1096), inverted for -n."""
    ref = _sel(parse_bed(texts[0], 3), chrom)
    u = merged_union([_sel(parse_bed(t, 3), chrom) for t in texts[1:]])
    out = []
    last_chrom = max(u) if u else None
    for r in ref:
        segs = u.get(r.chrom, [])
        ends = [e for _, e in segs]
        k = bisect.bisect_right(ends, r.start)
        ovl = 0.0
        while k < len(segs) and segs[k][0] < r.end:
            ovl += min(r.end, segs[k][1]) - max(r.start, segs[k][0])
            k += 1
        rng = float(r.end - r.start)
        is_elem = (ovl / rng >= thr) if use_pct else (ovl >= thr)
        if not use_pct and thr <= 0:
            # "-e 0" / "-n 0": 0 >= 0 holds for every row the loop reaches, but once the merged stream has no element
            # at or after the row (:1044-1049, :1053-1062) the row is reported as a non-element without the test
            is_elem = (bool(segs) and segs[-1][1] > r.start) or (last_chrom is not None and r.chrom < last_chrom)
        if is_elem != invert:
            out.append(echo_b3rest(r) + b"\n")
    return b"".join(out)


# --------------------------------------------------------------------------------------------------
# closest-features (declarative rule; see SURVEY.md 8c hazard 3 for where the streaming reference differs)
# --------------------------------------------------------------------------------------------------
def _cf_distance(c: Row, b: Row):
    """getDistance(c, b) (ClosestFeature.cpp:244-255): -inf / +inf across chromosomes (strcmp order), negative when c
    lies left of b, positive when right, 0 on overlap."""
    if c.chrom != b.chrom:
        return float("-inf") if c.chrom < b.chrom else float("inf")
    if c.end <= b.start:
        return -(b.start - c.end + 1)
    if b.end <= c.start:
        return c.start - b.end + 1
    return 0


def closest_pairs(refs: Sequence[Row], qrys: Sequence[Row], allow_overlaps: bool = True):
    """findDistances (ClosestFeature.cpp:260-413) restated statement by statement, including the push-back list
    (`read`, handed to BedReader::PushBack, BedReader.hpp:64-66: the list is re-read front first, ahead of whatever was
    pushed back earlier) -- the reference's answer is a function of this streaming state, not only of the two files:
    elements dropped ("delete") are gone for every later reference row, and elements held as left/right re-enter the
    stream after the ones pushed while they were held.  Returns [(left, right)] per reference row (Row or None)."""
    NEG, POS = float("-inf"), float("inf")
    cache: List[Row] = []          # BedReader::cache_: a stack, the back is read next
    pos = 0                        # next unread row of the query file

    def read_line():
        nonlocal pos
        if cache:
            return cache.pop()
        if pos < len(qrys):
            pos += 1
            return qrys[pos - 1]
        return None

    out = []
    for b in refs:
        left_dist, right_dist = NEG, POS
        left = right = None
        left_cached = False
        read: List[Row] = []
        c = None
        while True:
            c = read_line()
            if c is None:
                break
            dist = _cf_distance(c, b)
            if dist == NEG:                                   # :286-288 query is on an earlier chromosome: drop
                continue
            if dist == POS:                                   # :289-296 query is on a later chromosome
                if left is not None and not left_cached:
                    read.append(left)
                left_cached = left is not None
                if right is not None:
                    read.append(right)
                read.append(c)
                break
            if dist < 0 and dist >= left_dist:                # :300-309 new best left: everything pushed so far is dropped
                read.clear()
                left_dist, left, left_cached = dist, c, False
            elif dist < 0:                                    # :310-314
                if not left_cached:
                    read.append(left)
                left_cached = True
            elif dist > 0 and dist < right_dist:              # :315-323
                if left is not None and not left_cached:
                    read.append(left)
                left_cached = left is not None
                right_dist, right = dist, c
                read.append(c)
                break
            elif dist > 0:                                    # :324-331 read one too many
                if left is not None and not left_cached:
                    read.append(left)
                left_cached = left is not None
                if right is not None:
                    read.append(right)
                read.append(c)
                break
            elif allow_overlaps:                              # :332-388
                if c.start <= b.start:                        # hangs over the left edge
                    if left is not None and left.end <= c.end and not left_cached:
                        pass                                  # delete left
                    elif left is not None and not left_cached:
                        read.append(left)
                    left, left_dist, left_cached = c, 0, False
                elif b.end <= c.end:                          # hangs over the right edge
                    if left is not None and not left_cached:
                        read.append(left)
                    left_cached = left is not None
                    if right is not None:
                        read.append(right)
                    right, right_dist = c, 0
                else:                                         # contained in the reference row
                    centroid = (b.end - 1.0 + b.start) / 2.0  # getCentroid :236-239
                    prop = 0.0 if centroid < c.start else (centroid + 1 - c.start) / (c.end - c.start)  # :226-231
                    if left_dist == 0:
                        if prop < 0.5:
                            if not left_cached:
                                read.append(left)
                            left_cached = True
                            if right is not None:
                                read.append(right)
                            right, right_dist = c, 0
                        else:
                            if not left_cached:
                                read.append(left)
                            left_cached = True
                            read.append(c)
                    elif prop >= 0.5:                         # :369-378 new left: the push-back list is dropped
                        read.clear()
                        left_cached = False
                        left, left_dist = c, 0
                    else:
                        if left is not None and not left_cached:
                            read.append(left)
                        left_cached = left is not None
                        if right is not None:
                            read.append(right)
                        right, right_dist = c, 0
            else:                                             # :389-397 --no-overlaps: keep for later reference rows
                if left is not None and not left_cached:
                    read.append(left)
                    left_cached = True
                read.append(c)
        if c is None and left is not None and not left_cached:    # :403-406
            read.append(left)
        if c is None and right is not None:
            read.append(right)
        cache.extend(reversed(read))                          # PushBack(list): front of `read` is read first
        out.append((left, right))
    return out


def closest_features(ref_text: bytes, qry_text: bytes, dist: bool = False, closest: bool = False,
                     no_overlaps: bool = False, no_ref: bool = False, delim: bytes = b"|",
                     chrom: Optional[bytes] = None) -> bytes:
    """findDistances (applications/bed/closestfeats/src/ClosestFeature.cpp:260-413) + PrintAll/PrintShortest
    (Printers.hpp:46-205)."""
    refs = _sel(parse_bed(ref_text, 3), chrom)
    qrys = _sel(parse_bed(qry_text, 3), chrom)
    out = []
    for b, (left, right) in zip(refs, closest_pairs(refs, qrys, not no_overlaps)):
        ldist = _cf_distance(left, b) if left is not None else None
        rdist = _cf_distance(right, b) if right is not None else None
        parts: List[bytes] = []

        def show(x, d):
            parts.append(echo_b3rest(x) if x is not None else b"NA")
            if dist:
                parts.append(str(d).encode() if x is not None else b"NA")

        if not no_ref:
            parts.append(echo_b3rest(b))
        if closest:                                 # PrintShortest (Printers.hpp:120-205)
            if left is None and right is None:
                show(None, None)
            elif right is None or (left is not None and abs(ldist) <= abs(rdist)):
                show(left, ldist)
            else:
                show(right, rdist)
        else:
            show(left, ldist)
            show(right, rdist)
        out.append(delim.join(parts) + b"\n")
    return b"".join(out)


# ---------------------------------------------------------------------------------------------------------
# sort-bed (applications/bed/sort-bed/src/SortDetails.cpp:530-1208, Structures.hpp:47-76)
# ---------------------------------------------------------------------------------------------------------
class SortBedError(ValueError):
    """A line sort-bed rejects; .message is the text it prints on stderr (SortDetails.cpp:631-779, :833-853)."""

    def __init__(self, message: str):
        super().__init__(message)
        self.message = message


_SB_WS = b" \t\r\v\f\n"


def _sortbed_line(line: bytes, has_nl: bool, lineno: int, fname: str):
    """One data line (without its NL) -> (chrom, start, end, rest-or-None), as processData reads it."""
    def at(fmt):
        return SortBedError(fmt % (lineno, fname))
    if line[:1] in (b" ", b"\t"):
        raise at("Row begins with a tab or space at line %d in %s.\n")
    def find_sep(pos):
        k = pos
        while k < len(line) and line[k:k + 1] not in (b"\t", b" "):
            k += 1
        return k
    c = find_sep(0)
    if c == len(line):
        raise at("No tabs/spaces found at line %d in %s.\n")                                    # :657-663
    if c > 127:
        raise at("Chromosome name too long at line %d in %s.\n"
                 "Check that you have unix newlines (cat -A) or increase TOKEN_CHR_MAX_LENGTH in BEDOPS.Constants.hpp "
                 "and recompile BEDOPS.\n")
    d = find_sep(c + 1)
    if d == len(line):
        raise at("No tabs/spaces found after the start coordinate (or no start coordinate at all) at line %d in %s.\n")
    if d - (c + 1) > 12:
        raise at("Start coordinate is too large.  Max decimal digits allowed is 12 in BEDOPS.Constants.hpp.  See line %d in %s.\n")
    if d == c + 1:
        raise at("Consecutive tabs and/or spaces between chromosome and start coordinate.  See line %d in %s.\n")
    if not line[c + 1:d].isdigit():
        raise at("Non-numeric start coordinate.  See line %d in %s.\n(remember that chromosome names should not contain spaces.)\n")
    e = find_sep(d + 1)
    if e == len(line) and not has_nl:                                                           # :719-730
        raise at("No end of line found at %d in %s.\nMay need to increase BED_LINE_LEN and recompile.\n"
                 "First check that you have unix newlines (cat -A).")
    if e - (d + 1) > 12:
        raise at("End coordinate is too large.  Max decimal digits allowed is 12 in BEDOPS.Constants.hpp.  See line %d in %s.\n")
    if e == d + 1:
        raise at("Extra tab and/or space found in between start and end coordinates.  See line %d in %s.\n")
    if not line[d + 1:e].isdigit():
        raise at("Non-numeric end coordinate.  See line %d in %s.\n")
    start, end = int(line[c + 1:d]), int(line[d + 1:e])
    if end <= start:
        raise at("Error on line %d in %s. Genomic end coordinate is less than (or equal to) start coordinate.\n")
    q = e                                                                                       # sscanf "\t%[^\n]s\n", :764
    while q < len(line) and line[q:q + 1] in (b" ", b"\t", b"\r", b"\v", b"\f"):
        q += 1
    rest = line[q:] if q < len(line) else None
    if rest is not None:
        k = 0
        while k < len(rest) and rest[k:k + 1] not in (b"\t", b" "):
            k += 1
        if k > 16383:
            raise at("ID field too long at line %d in %s.\nCheck that you have unix newlines (cat -A) or increase "
                     "TOKEN_ID_MAX_LENGTH in BEDOPS.Constants.hpp and recompile BEDOPS.\nYou may instead choose to put a "
                     "dummy id column (like 'id') in as the 4th field to fix this.\n")
    return line[:c], start, end, rest


def sort_bed(texts: Sequence[bytes], names: Optional[Sequence[str]] = None) -> bytes:
    """sort-bed file1 file2 ...: rows of all files ordered by chromosome (strcmp), start, end, rest (strcmp; no rest
    first, bcd_cmp Structures.hpp:47-76), printed "%s\\t%ld\\t%ld[\\t%s]\\n" (printBed :1120-1140).  Empty lines are
    skipped (:625-629); browser/track/#/@ lines are skipped until the first data row of each file (:645-653)."""
    rows = []
    for fi, text in enumerate(texts):
        fname = names[fi] if names else "file%d" % fi
        head = True
        lines = text.split(b"\n")
        terminated = text.endswith(b"\n")
        if terminated:
            lines.pop()
        for li, line in enumerate(lines):
            if line == b"":
                continue
            has_nl = terminated or li + 1 < len(lines)
            if line[:1] not in (b" ", b"\t") and head and (line.startswith(b"browser") or line.startswith(b"track")
                                                            or line.startswith(b"#") or line.startswith(b"@")):
                continue
            rows.append(_sortbed_line(line, has_nl, li + 1, fname))
            head = False
    # bytes compare like strcmp (unsigned bytes); None (no rest) sorts first
    rows.sort(key=lambda r: (r[0], r[1], r[2], (0, b"") if r[3] is None else (1, r[3])))
    out = []
    for chrom, s, e, rest in rows:
        out.append(chrom + b"\t%d\t%d" % (s, e) + (b"\t" + rest if rest is not None else b"") + b"\n")
    return b"".join(out)


# ---------------------------------------------------------------------------------------------------------
# bedops --range L:R (BedPadReader.hpp:116-277) and --partition (Bedops.cpp:615-653, :1249-1335)
# ---------------------------------------------------------------------------------------------------------
U64 = 1 << 64


def pad_rows(rows: Sequence[Row], lpad: int, rpad: int) -> List[Row]:
    """BedPadReader::ReadLine over one file: start += lpad, end += rpad.
    * rpad < 0 or lpad > 0 (:133-141): a row whose padded end does not pass its padded start vaporises; the start is an
      unsigned 64-bit value, so start + lpad < 0 wraps to a huge number and the row vaporises too -- except in the zone
      the constructor's getFirst() (:194-277) consumed when lpad < 0: from the first row of the file up to the first
      surviving row with start > |lpad|, where starts are clamped to 0 instead.
    * lpad < 0 otherwise (:142-155): at every chromosome getFirst() clamps the starts <= |lpad| to 0 and re-sorts those
      rows by their end (input order on ties); the others shift.
    * rpad > 0 (:156-160): ends shift."""
    import copy
    out: List[Row] = []
    if lpad == 0 and rpad == 0:
        return list(rows)
    lpd = abs(lpad)

    def clamp_group(group):          # rows whose start became 0: ordered by end, stable (multiset + tie list, :203-275)
        return sorted(group, key=lambda r: (r.chrom, r.start, r.end))

    def mk(r, s, e):
        q = copy.copy(r)
        q.start, q.end = s, e
        return q

    if rpad < 0 or lpad > 0:
        k = 0
        if lpad < 0:                 # the constructor's getFirst()
            zone = []
            while k < len(rows):
                r = rows[k]
                k += 1
                if r.start > lpd:
                    s = r.start - lpd
                    if (r.end + rpad) % U64 > s:      # unsigned arithmetic here (:209): an end below |rpad| wraps and is kept
                        zone.append(mk(r, s, (r.end + rpad) % U64))
                        break
                    continue
                if r.end + rpad <= 0:
                    continue
                zone.append(mk(r, 0, r.end + rpad))
            out.extend(clamp_group(zone))
        for r in rows[k:]:
            s = (r.start + lpad) % U64
            if float(r.end) + rpad > s:
                out.append(mk(r, s, r.end + rpad))
        return out
    if lpad < 0:
        k = 0
        last = None
        while k < len(rows):
            r = rows[k]
            if r.chrom != last:
                zone = []
                while k < len(rows):
                    r = rows[k]
                    k += 1
                    if r.start > lpd:
                        zone.append(mk(r, r.start - lpd, r.end + rpad))
                        break
                    zone.append(mk(r, 0, r.end + rpad))
                zone = clamp_group(zone)
                out.extend(zone)
                last = zone[-1].chrom
                # rows handed out of the cache set lastChr_ as they go (:125-127): the last one names the chromosome
            else:
                out.append(mk(r, r.start - lpd, r.end + rpad))
                k += 1
        return out
    return [mk(r, r.start, r.end + rpad) for r in rows]


def merge_everything(files: Sequence[Sequence[Row]]) -> bytes:
    """nextUnionAllLine (Bedops.cpp:1468-1516) as the k-way merge it is: the smallest head by (chromosome, start, end,
    rest), strict comparisons, so the lower-numbered file wins ties.  On sort-bed-sorted files this equals a full sort; on
    padded files, where rows clamped to start 0 keep their input order, it does not."""
    heads = [0] * len(files)
    out = []
    while True:
        best = -1
        for i, f in enumerate(files):
            if heads[i] >= len(f):
                continue
            if best < 0:
                best = i
                continue
            a, b = f[heads[i]], files[best][heads[best]]
            if (a.chrom, a.start, a.end, a.rest3) < (b.chrom, b.start, b.end, b.rest3):
                best = i
        if best < 0:
            break
        out.append(echo_b3rest(files[best][heads[best]]) + b"\n")
        heads[best] += 1
    return b"".join(out)


def _rows_to_text(rows: Sequence[Row]) -> bytes:
    return b"".join(echo_b3rest(r) + b"\n" for r in rows)


def pad_text(text: bytes, lpad: int, rpad: int) -> bytes:
    """a file as the set operators see it under --range L:R"""
    return _rows_to_text(pad_rows(parse_bed(text, 3), lpad, rpad))


def bedops_partition(texts: Sequence[bytes], chrom: Optional[bytes] = None) -> bytes:
    """doPartitions: the pieces between consecutive break points (all starts and ends of all files) of a chromosome that
    some row covers; touching rows are not joined, duplicate rows give one piece."""
    per: Dict[bytes, List[Tuple[int, int]]] = {}
    for t in texts:
        for r in _sel(parse_bed(t, 3), chrom):
            per.setdefault(r.chrom, []).append((r.start, r.end))
    out = []
    for c in sorted(per):
        iv = per[c]
        pts = sorted({p for s, e in iv for p in (s, e)})
        iv.sort()
        # coverage by a sweep over the break points
        import heapq
        ends: List[int] = []
        k = 0
        for a, b in zip(pts, pts[1:]):
            while k < len(iv) and iv[k][0] <= a:
                heapq.heappush(ends, iv[k][1])
                k += 1
            while ends and ends[0] <= a:
                heapq.heappop(ends)
            if ends:
                out.append(c + b"\t%d\t%d\n" % (a, b))
    return b"".join(out)


# ---------------------------------------------------------------------------------------------------------
# Starch v2 archives (interfaces/src/data/starch/unstarchHelpers.c)
# ---------------------------------------------------------------------------------------------------------
def starch_streams(archive: bytes):
    """[(chromosome, delta-coded text)] of a Starch v2 archive: magic ca5cade5, the per-chromosome bzip2/gzip streams, the
    JSON metadata and the 127-byte footer holding its offset (starchHelpers.c STARCH2_writeStarchFooter)."""
    import bz2
    import json
    import zlib
    assert archive[:4] == b"\xca\x5c\xad\xe5", "not a Starch v2 archive"
    md_off = int(archive[-127:-107])                     # the footer on disk is STARCH2_MD_FOOTER_LENGTH - 1 = 127 bytes
    md = json.loads(archive[md_off:-127])                # (starchMetadataHelpers.c:1113-1119, :1178)
    comp = md["archive"]["compressionFormat"]            # 0 bzip2, 1 gzip (starchMetadataHelpers.h CompressionType)
    out, at = [], 4
    for st in md["streams"]:
        n = int(st["size"])
        blob = archive[at:at + n]
        at += n
        text = bz2.decompress(blob) if comp == 0 else zlib.decompress(blob, 15 + 32)
        out.append((st["chromosome"].encode(), text))
    return out


def unstarch(archive: bytes, chrom: Optional[bytes] = None) -> bytes:
    """UNSTARCH_reverseTransformHeaderlessInput (unstarchHelpers.c:1161-1238): "p<len>" sets the length of the rows that
    follow; "<d>[\\t<rest>]" is a row with start = end of the previous row + d, end = start + len."""
    out = []
    for c, text in starch_streams(archive):
        if chrom is not None and chrom != b"all" and c != chrom:
            continue
        last_end = plen = 0
        for line in text.split(b"\n"):
            if not line:
                continue
            if line[:1] == b"p":
                plen = int(line[1:])
                continue
            d, _, rest = line.partition(b"\t")
            start = last_end + int(d)
            last_end = start + plen
            out.append(c + b"\t%d\t%d" % (start, last_end) + (b"\t" + rest if rest else b"") + b"\n")
    return b"".join(out)
