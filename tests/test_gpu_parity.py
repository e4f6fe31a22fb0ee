"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
  * the reference's own golden vectors (TestPlan.xml in-scope orders, docs worked examples),
  * the Python oracle on seeded synthetic inputs (and the committed hashes of the reference binaries' output),
  * the reference binaries themselves (oracle/_ref/bin) through the drop-in command lines,
plus parser column parity and edge cases (empty / ragged / unterminated input, chromosome handling).
Bar: byte-identical output."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import REFBIN, have_ref, load_golden
import bed_oracle as O
import oracle_cli

pytestmark = pytest.mark.gpu

TESTPLAN = load_golden("testplan.json")
DOCS = load_golden("docs.json")
SYN = load_golden("synthetic.json")


@pytest.fixture(scope="module")
def kit():
    import bedops_b200
    k = bedops_b200.BedKit(0)
    yield k
    k.close()


def first_diff(a: bytes, b: bytes) -> str:
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return "line %d: got %r expected %r (lens %d/%d lines)" % (i + 1, x[:200], y[:200], len(la), len(lb))
    return "length differs: %d vs %d lines" % (len(la), len(lb))


def assert_same(got: bytes, exp: bytes):
    assert got == exp, first_diff(got, exp)


# ---- reader ------------------------------------------------------------------------------------------------
def test_parser_columns_match_oracle(kit, synth_files):
    from bedops_b200._lib import COL_LINE, COL_SCORE
    text = synth_files["m.bed"]
    bed = kit.load(text, 5, COL_SCORE | COL_LINE)
    rows = O.parse_bed(text, 5)
    st, en, sc, lo = bed.columns(score=True, line=True)
    assert bed.rows == len(rows)
    assert np.array_equal(st, np.array([r.start for r in rows], dtype=np.uint32))
    assert np.array_equal(en, np.array([r.end for r in rows], dtype=np.uint32))
    assert np.array_equal(sc, np.array([r.score for r in rows], dtype=np.float64))
    # line offsets point at the chromosome token of each row
    offs = np.cumsum([0] + [len(l) + 1 for l in text.split(b"\n")[:-1]])[:-1]
    assert np.array_equal(lo, offs.astype(np.uint64))
    names = [c for c, _ in bed.chroms()]
    assert names == sorted({r.chrom.decode() for r in rows})
    assert sum(n for _, n in bed.chroms()) == len(rows)
    bed.free()


def test_parser_float_scores_are_strtod_exact(kit):
    from bedops_b200 import synth
    from bedops_b200._lib import COL_SCORE
    text = synth.bed_text(20000, 7, fields=5, float_scores=True)
    extra = b"chrZ\t1\t5\tx\t1e3\nchrZ\t2\t6\tx\t-0.000001\nchrZ\t3\t7\tx\t123456789012345.678\nchrZ\t4\t8\tx\t+.5\nchrZ\t5\t9\tx\t7.\n"
    text += extra
    bed = kit.load(text, 5, COL_SCORE)
    _, _, sc, _ = bed.columns(score=True)
    exp = np.array([r.score for r in O.parse_bed(text, 5)], dtype=np.float64)
    assert np.array_equal(sc.view(np.uint64), exp.view(np.uint64))
    bed.free()


def test_parser_edge_cases(kit):
    # empty input, blank lines, CRLF, unterminated last line (dropped), spaces as separators, leading zeros, '+'
    assert kit.load(b"", 3).rows == 0
    assert kit.load(b"\n\n\n", 3).rows == 0
    assert kit.load(b"chr1\t1\t2", 3).rows == 0
    t = b"\nchr1 \t 0007  +12\textra stuff\r\n\n  \nchr1\t20\t30\nchr2\t5\t6\nchr2\t7\t9"
    bed = kit.load(t, 3, 1)
    st, en, _, lo = bed.columns(line=True)
    assert bed.rows == 3 and st.tolist() == [7, 20, 5] and en.tolist() == [12, 30, 6]
    assert bed.chroms() == [("chr1", 2), ("chr2", 1)]
    out = kit.setop("not-element-of", [bed, kit.load(b"chr9\t1\t2\n", 3)], 1.0, False)
    assert out == b"chr1\t7\t12\textra stuff\r\nchr1\t20\t30\nchr2\t5\t6\n"
    assert out == O.bedops_element_of([t, b"chr9\t1\t2\n"], 1.0, False, True)


def test_parser_rejects_what_it_cannot_parse_exactly(kit):
    import bedops_b200
    for bad, code in ((b"chr1\tx\t5\n", 4), (b"chr1\t1\n", 4), (b"chr1\t1\t4294967295\n", 5),
                      (b"chr2\t1\t2\nchr1\t1\t2\n", 8)):
        with pytest.raises(bedops_b200.BedKitError) as e:
            kit.load(bad, 3)
        assert e.value.code == code, bad
    with pytest.raises(bedops_b200.BedKitError) as e:
        kit.load(b"chr1\t1\t2\tid\n", 5, 2)   # B5Rest needs a score column
    assert e.value.code == 4
    with pytest.raises(bedops_b200.BedKitError) as e:
        kit.load(b"\xca\x5c\xad\xe5\x00\x00", 3)   # starch magic
    assert e.value.code == 7


def test_tile_boundaries_and_long_lines(kit):
    # lines that straddle the 8 KiB parser tiles, a line longer than a tile, many chromosomes
    rng = np.random.default_rng(5)
    lines = []
    pos = 0
    for c in range(40):
        pos = 0
        for k in range(rng.integers(1, 400)):
            pos += int(rng.integers(1, 50))
            rest = b"" if k % 3 else b"\t" + b"x" * int(rng.integers(0, 300))
            if k == 7 and c == 3:
                rest = b"\t" + b"y" * 20000
            lines.append(b"c%03d\t%d\t%d%s" % (c, pos, pos + int(rng.integers(1, 100)), rest))
    text = b"\n".join(lines) + b"\n"
    bed = kit.load(text, 3, 1)
    rows = O.parse_bed(text, 3)
    st, en, _, _ = bed.columns()
    assert st.tolist() == [r.start for r in rows] and en.tolist() == [r.end for r in rows]
    assert [c for c, _ in bed.chroms()] == ["c%03d" % c for c in range(40)]
    assert_same(kit.bedmap(bed, None, ["echo", "count", "bases"]), O.bedmap(text, None, ["echo", "count", "bases"]))
    assert_same(kit.setop("merge", [bed]), O.bedops_merge([text]))


# ---- golden vectors of the reference ------------------------------------------------------------------------------
@pytest.mark.parametrize("case", TESTPLAN, ids=lambda c: "order%d" % c["order"])
def test_testplan_golden(kit, case):
    files = {k: v.encode() for k, v in case["files"].items()}
    got = oracle_cli.run_kit(kit, "bedops", case["argv"], files)
    assert_same(got, case["raw_stdout"].encode())


@pytest.mark.parametrize("ex", DOCS["examples"], ids=lambda e: e["source"].split()[0])
def test_docs_examples(kit, ex):
    files = {k: v.encode() for k, v in DOCS["files"].items()}
    stdin = ex["stdin"].encode() if ex["stdin"] else None
    assert_same(oracle_cli.run_kit(kit, ex["tool"], ex["argv"], files, stdin), ex["expected"].encode())


# ---- synthetic inputs: CUDA vs oracle vs hashes of the reference binaries' output ------------------------------------
@pytest.mark.parametrize("case", SYN["cases"],
                         ids=lambda c: c["tool"] + "_" + "_".join(c["argv"]).replace("\t", "TAB"))
def test_synthetic_vs_reference_hash(kit, case, synth_files):
    got = oracle_cli.run_kit(kit, case["tool"], case["argv"], synth_files)
    assert len(got) == case["nbytes"]
    assert hashlib.sha256(got).hexdigest() == case["sha256"]


def test_synthetic_vs_oracle_bytes(kit, synth_files):
    for tool, argv in (("bedops", ["-m", "m.bed", "m2.bed"]), ("bedops", ["-i", "m.bed", "m2.bed", "r.bed"]),
                       ("bedops", ["-e", "30%", "m.bed", "m2.bed"]), ("bedops", ["-n", "10", "m.bed", "r.bed"]),
                       ("bedmap", ["--echo", "--count", "--mean", "--bases", "--sum", "--max", "--min", "r.bed", "m.bed"]),
                       ("bedmap", ["--prec", "0", "--mean", "--echo-ref-size", "--echo-ref-name", "--echo-ref-row-id", "r.bed", "m.bed"]),
                       ("bedmap", ["--prec", "12", "--mean", "r.bed", "m.bed"]),
                       ("bedmap", ["--sci", "--mean", "--sum", "--max", "r.bed", "m.bed"]),
                       ("bedmap", ["--sci", "--prec", "17", "--mean", "--min", "r.bed", "m.bed"]),
                       ("bedmap", ["--sci", "--prec", "0", "--mean", "r.bed", "m.bed"]),
                       ("bedmap", ["--multidelim", "::", "--delim", "\t", "--echo-map-id", "--count", "r.bed", "u.bed"])):
        assert_same(oracle_cli.run_kit(kit, tool, argv, synth_files), oracle_cli.run(tool, argv, synth_files))


def test_dense_map_file_takes_the_block_skipping_scan(kit, synth_files, monkeypatch):
    """>= 32 map rows per reference row: both window kernels skip 32-row blocks in which no end reaches the reference rows
    (k_map_group: whole chunks, when all four groups of the warp agree; k_map_stats<SKIP> under BEDKIT_MAP_KERNEL=row)."""
    files = dict(synth_files)
    files["few.bed"] = b"".join(l + b"\n" for l in synth_files["dr.bed"].split(b"\n")[:-1][::12])
    assert files["dm.bed"].count(b"\n") >= 32 * files["few.bed"].count(b"\n")
    for env in (None, "row"):
        if env:
            monkeypatch.setenv("BEDKIT_MAP_KERNEL", env)
        else:
            monkeypatch.delenv("BEDKIT_MAP_KERNEL", raising=False)
        for argv in (["--echo", "--count", "--mean", "--bases", "few.bed", "dm.bed"],
                     ["--count", "--sum", "--max", "--min", "--indicator", "few.bed", "dm.bed"],
                     ["--bp-ovr", "50", "--count", "--bases", "few.bed", "dm.bed"]):
            assert_same(oracle_cli.run_kit(kit, "bedmap", argv, files), oracle_cli.run("bedmap", argv, files))


def test_everything_orders_ties_by_the_rest_of_the_line_then_by_file(kit):
    """bedops -u: identical coordinates are ordered by strcmp of the rest of the line, full ties by file number
    (nextUnionAllLine, Bedops.cpp:1468-1516); BED3 and BED5 rows mixed, chromosomes missing from some files."""
    files = {
        "t1.bed": b"chr1\t10\t20\tb\t1\nchr1\t10\t20\tb\t1\nchr1\t10\t30\nchr1\t50\t60\tz\nchr2\t5\t9\tq\n",
        "t2.bed": b"chr1\t10\t20\ta\t9\nchr1\t10\t20\tb\t1\nchr1\t10\t30\tx\nchr1\t50\t60\nchr3\t1\t2\n",
        "t3.bed": b"chr1\t10\t20\nchr1\t10\t20\tb\nchr1\t50\t60\ty\nchr2\t5\t9\tq\n",
    }
    for argv in (["-u", "t1.bed", "t2.bed", "t3.bed"], ["-u", "t3.bed", "t2.bed", "t1.bed"], ["-u", "t2.bed"],
                 ["--chrom", "chr2", "-u", "t1.bed", "t3.bed"]):
        assert_same(oracle_cli.run_kit(kit, "bedops", argv, files), oracle_cli.run("bedops", argv, files))


def test_nested_and_duplicate_intervals(kit):
    # adversarial nesting: one chromosome-long interval, duplicates, touching and abutting intervals
    m = [b"chr1\t0\t1000000\tbig\t5"]
    for k in range(2000):
        s = 10 + k * 400
        m.append(b"chr1\t%d\t%d\ta%d\t%d" % (s, s + 150, k, k % 17))
        if k % 5 == 0:
            m.append(b"chr1\t%d\t%d\tb%d\t%d" % (s, s + 150, k, k % 11))
        if k % 7 == 0:
            m.append(b"chr1\t%d\t%d\tc%d\t%d" % (s + 150, s + 300, k, 3))
    m.sort(key=lambda l: (int(l.split(b"\t")[1]), int(l.split(b"\t")[2])))
    mt = b"\n".join(m) + b"\n"
    r = [b"chr1\t%d\t%d\tr%d" % (k * 333, k * 333 + 1 + (k * 37) % 900, k) for k in range(2500)]
    rt = b"\n".join(r) + b"\n"
    files = {"r.bed": rt, "m.bed": mt}
    for tool, argv in (("bedmap", ["--echo", "--count", "--bases", "--mean", "--max", "--min", "r.bed", "m.bed"]),
                       ("bedmap", ["--count", "--bases", "m.bed"]), ("bedmap", ["--range", "200", "--count", "r.bed", "m.bed"]),
                       ("bedops", ["-m", "m.bed", "r.bed"]), ("bedops", ["-i", "m.bed", "r.bed"]),
                       ("bedops", ["-e", "100%", "r.bed", "m.bed"]), ("bedops", ["-n", "m.bed", "r.bed"])):
        assert_same(oracle_cli.run_kit(kit, tool, argv, files), oracle_cli.run(tool, argv, files))


def test_single_file_mode_echo_uses_the_map_record_type(kit, synth_files):
    """bedmap with one file: the reference row is a B4Rest/B5Rest, so --echo re-prints the id and, with a score
    operation, column 5 as "%lf" (Bedmap.cpp:676-700, Bed.hpp:740-743)."""
    t = b"chr1\t10\t20\tz\t1\textra\nchr1\t15\t30\ty\t2.5\nchr2\t5\t9\tx\t-3e2\tq\n"
    exp = {("echo", "mean"): b"chr1\t10\t20\tz\t1.000000\textra|1.750000\nchr1\t15\t30\ty\t2.500000|1.750000\nchr2\t5\t9\tx\t-300.000000\tq|-300.000000\n",
           ("echo", "echo-map-id"): b"chr1\t10\t20\tz\t1\textra|z;y\nchr1\t15\t30\ty\t2.5|z;y\nchr2\t5\t9\tx\t-3e2\tq|x\n",
           ("echo", "count"): b"chr1\t10\t20\tz\t1\textra|2\nchr1\t15\t30\ty\t2.5|2\nchr2\t5\t9\tx\t-3e2\tq|1\n"}
    for ops, want in exp.items():   # `want` is the reference binary's output for this input
        argv = ["--" + o for o in ops] + ["s.bed"]
        assert_same(oracle_cli.run("bedmap", argv, {"s.bed": t}), want)
        assert_same(oracle_cli.run_kit(kit, "bedmap", argv, {"s.bed": t}), want)
    argv = ["--echo", "--sum", "--count", "m.bed"]
    assert_same(oracle_cli.run_kit(kit, "bedmap", argv, synth_files), oracle_cli.run("bedmap", argv, synth_files))


def test_closest_features_vs_oracle(kit, synth_files):
    """Declarative closest-features rule (non-nested reference rows: the case where the reference's streaming state
    machine and the rule agree, SURVEY 8c hazard 3), every output mode, plus nested QUERY rows and --chrom."""
    for argv in (["r.bed", "m.bed"], ["--dist", "r.bed", "m.bed"], ["--closest", "--dist", "r.bed", "m.bed"],
                 ["--no-ref", "--delim", "\t", "r.bed", "m3.bed"], ["--chrom", "chr8", "--dist", "r.bed", "m.bed"],
                 ["--dist", "r.bed", "r.bed"]):
        assert_same(oracle_cli.run_kit(kit, "closest-features", argv, synth_files),
                    oracle_cli.run("closest-features", argv, synth_files))
    # hand-made: ties on the left end (later row wins), containment by centroid, touching neighbours, no neighbours
    q = b"chr1\t5\t10\ta\nchr1\t7\t10\tb\nchr1\t20\t30\tc\nchr1\t22\t24\td\nchr1\t26\t28\te\nchr1\t40\t50\tf\nchr3\t1\t2\tg\n"
    r = b"chr1\t12\t15\tr1\nchr1\t21\t29\tr2\nchr1\t30\t40\tr3\nchr1\t60\t70\tr4\nchr2\t1\t5\tr5\nchr3\t0\t1\tr6\n"
    files = {"q.bed": q, "r.bed": r}
    for argv in (["--dist", "r.bed", "q.bed"], ["--closest", "r.bed", "q.bed"], ["--dist", "q.bed", "q.bed"]):
        assert_same(oracle_cli.run_kit(kit, "closest-features", argv, files), oracle_cli.run("closest-features", argv, files))


def test_float_scores_within_tolerance(kit):
    """--mean/--sum on floating-point scores: <= 1e-12 relative to the exactly rounded per-row sum (math.fsum),
    before --prec formatting (north_star).  Compared at --prec 15 through the printed text."""
    import math
    from bedops_b200 import synth
    m = synth.bed_text(30000, 11, fields=5, float_scores=True)
    r = synth.bed_text(3000, 12, synth.REF_SHAPE, fields=5)
    got = oracle_cli.run_kit(kit, "bedmap", ["--prec", "15", "--sum", "--mean", "r.bed", "m.bed"], {"r.bed": r, "m.bed": m})
    refs, maps = O.parse_bed(r, 3), O.by_chrom(O.parse_bed(m, 5))
    lines = got.split(b"\n")[:-1]
    assert len(lines) == len(refs)
    worst = 0.0
    for ref, line in zip(refs, lines):
        hits = [x.score for x in maps.get(ref.chrom, []) if x.start < ref.end and x.end > ref.start]
        a, b = line.split(b"|")
        if not hits:
            assert a == b == b"NAN"
            continue
        exact, scale = math.fsum(hits), math.fsum(abs(h) for h in hits)
        worst = max(worst, abs(float(a) - exact) / scale, abs(float(b) - exact / len(hits)) / scale * len(hits))
    assert worst <= 1e-12, worst


# ---- the drop-in command lines against the reference binaries ------------------------------------------------------
def test_pipelined_host_call_equals_the_unsplit_call(kit, synth_files):
    """bk_bedmap_host (chromosome groups, transfers overlapped with kernels) == bk_load_bed x2 + bk_bedmap, including
    chromosomes that exist in only one of the files."""
    from bedops_b200._lib import COL_ID, COL_LINE, COL_SCORE
    ref, mp = synth_files["r.bed"], synth_files["m.bed"]
    extra_ref = b"chrZ\t5\t50\tonly_in_ref\t1\n"
    mp_gap = b"".join(l + b"\n" for l in mp.split(b"\n")[:-1] if not l.startswith(b"chr11\t")) + b"chrZZ\t1\t9\tx\t2\n"
    for r, m in ((ref, mp), (ref + extra_ref, mp_gap)):
        rn, mn = np.frombuffer(r, dtype=np.uint8), np.frombuffer(m, dtype=np.uint8)
        for ops, mf, mc in ((["echo", "count", "mean", "bases"], 5, COL_SCORE),
                            (["count", "echo-map-id"], 4, COL_ID | COL_LINE)):
            rb, mb = kit.load(r, 3, COL_LINE), kit.load(m, mf, mc)
            exp = kit.bedmap(rb, mb, ops)
            got = kit.bedmap_host(rn.ctypes.data, len(r), 3, COL_LINE, mn.ctypes.data, len(m), mf, mc, ops)
            assert_same(got, exp)
            rb.free()
            mb.free()
    # errors come from the unsplit path: same message, row numbers of the whole file
    bad = ref + b"chrZ\t10\tnot_a_number\n"
    bn = np.frombuffer(bad, dtype=np.uint8)
    mn = np.frombuffer(mp, dtype=np.uint8)
    with pytest.raises(Exception) as e1:
        kit.bedmap_host(bn.ctypes.data, len(bad), 3, COL_LINE, mn.ctypes.data, len(mp), 5, COL_SCORE, ["count"])
    with pytest.raises(Exception) as e2:
        kit.load(bad, 3, COL_LINE)
    assert str(e1.value) == str(e2.value)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_cli_tools_byte_identical_to_reference_binaries(tmp_path, synth_files):
    import bedops_b200
    for n, c in synth_files.items():
        (tmp_path / n).write_bytes(c)
    runs = [("bedops", ["-m", "m.bed", "r.bed"]), ("bedops", ["--ec", "-i", "m.bed", "m2.bed"]),
            ("bedops", ["-e", "50%", "r.bed", "m.bed"]), ("bedops", ["--chrom", "chr5", "-n", "1", "r.bed", "m.bed"]),
            ("bedmap", ["--echo", "--count", "--mean", "--bases", "r.bed", "m.bed"]),
            ("bedmap", ["--faster", "--delim", "\\t", "--sum", "--max", "--echo-map-id", "r.bed", "u.bed"]),
            ("bedmap", ["--count", "m3.bed"]), ("closest-features", ["--dist", "r.bed", "m.bed"]),
            ("bedops", ["-u", "r.bed", "m3.bed", "m.bed"]), ("bedops", ["-c", "-L", "r.bed", "m3.bed"]),
            ("bedops", ["-w", "500", "--stagger", "200", "-x", "r.bed", "m3.bed"]), ("bedops", ["--chop", "r.bed"]),
            ("bedops", ["-d", "m.bed", "r.bed"]), ("bedops", ["-s", "m.bed", "r.bed", "m2.bed"]),
            ("bedmap", ["--echo", "--echo-map", "--echo-map-score", "--bases-uniq-f", "dr.bed", "dm.bed"]),
            ("bedmap", ["--range", "200", "--echo-map-range", "--echo-map-size", "--echo-overlap-size", "--bases-uniq", "dr.bed", "dm.bed"]),
            ("closest-features", ["--closest", "--delim", "\\t", "r.bed", "m3.bed"])]
    for tool, argv in runs:
        ours = subprocess.run([bedops_b200.tool_path(tool)] + argv, cwd=tmp_path, capture_output=True)
        ref = subprocess.run([os.path.join(REFBIN, tool)] + argv, cwd=tmp_path, capture_output=True)
        assert ours.returncode == ref.returncode == 0, ours.stderr
        assert_same(ours.stdout, ref.stdout)
    # stdin as the reference file
    ours = subprocess.run([bedops_b200.tool_path("bedmap"), "--echo", "--indicator", "-", "m.bed"], cwd=tmp_path,
                          input=synth_files["r.bed"], capture_output=True)
    ref = subprocess.run([os.path.join(REFBIN, "bedmap"), "--echo", "--indicator", "-", "m.bed"], cwd=tmp_path,
                         input=synth_files["r.bed"], capture_output=True)
    assert_same(ours.stdout, ref.stdout)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_cli_sharded_mode_is_byte_identical(tmp_path, synth_files):
    """BEDKIT_GPUS=N: contiguous chromosome groups, one host thread + ctx per shard, outputs concatenated in shard
    order.  On a single-GPU box the shards share device 0 (BEDKIT_SHARE_DEVICE); on a multi-GPU box they spread."""
    import bedops_b200
    import torch
    for n, c in synth_files.items():
        (tmp_path / n).write_bytes(c)
    env = dict(os.environ, BEDKIT_GPUS="3")
    if torch.cuda.device_count() < 3:
        env["BEDKIT_SHARE_DEVICE"] = "1"
    for tool, argv in (("bedmap", ["--echo", "--count", "--mean", "--bases", "r.bed", "m.bed"]),
                       ("bedops", ["-m", "m.bed", "m2.bed", "r.bed"]), ("bedops", ["-e", "1", "r.bed", "m.bed"]),
                       ("closest-features", ["--dist", "r.bed", "m.bed"])):
        ours = subprocess.run([bedops_b200.tool_path(tool)] + argv, cwd=tmp_path, capture_output=True, env=env)
        ref = subprocess.run([os.path.join(REFBIN, tool)] + argv, cwd=tmp_path, capture_output=True)
        assert ours.returncode == 0, ours.stderr
        assert_same(ours.stdout, ref.stdout)


EC_CASES = {
    "unsorted_start.bed": b"chr1\t10\t20\nchr1\t5\t9\n",
    "unsorted_end.bed": b"chr1\t10\t20\tb\nchr1\t10\t15\ta\n",
    "unsorted_rest.bed": b"chr1\t10\t20\tb\t1\nchr1\t10\t20\ta\t1\n",
    "unsorted_chrom.bed": b"chr2\t10\t20\nchr1\t5\t9\n",
    "space_chrom.bed": b"track name=x\n#c\nchr1\t10\t20\tid\t3\nchr1 \t12\t30\n",
    "space_start.bed": b"chr1\t1 0\t20\n",
    "nonnum_end.bed": b"chr1\t10\t2x0\tid\t3\n",
    "neg_start.bed": b"chr1\t-10\t20\n",
    "two_tabs.bed": b"chr1\t\t20\t30\n",
    "no_tabs.bed": b"chr1\n",
    "empty_line.bed": b"chr1\t1\t2\n\nchr1\t3\t4\n",
    "end_le_start.bed": b"chr1\t10\t20\tid\t1\nchr1\t30\t30\tid\t1\n",
    "too_many_digits.bed": b"chr1\t1234567890123\t1234567890124\n",
    "short_cols.bed": b"chr1\t10\t20\tid\t3\nchr1\t12\t30\tid\n",
    "bad_score.bed": b"chr1\t10\t20\tid\t3.4.5\n",
    "bad_score2.bed": b"chr1\t10\t20\tid\t3-\n",
    "nested.bed": b"chr1\t10\t100\tid\t1\nchr1\t20\t30\tid\t2\n",
    "headers_ok.bed": b"browser x\ntrack y\n@hd\n#c\nchr1\t10\t20\tid\t3\nchr1\t12\t30\tid\t5",
}


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_ec_validation_matches_reference_messages(tmp_path):
    """--ec: same stdout, same stderr text (message + row number), same exit code as the reference tools."""
    import bedops_b200
    for n, c in EC_CASES.items():
        (tmp_path / n).write_bytes(c)
    (tmp_path / "good.bed").write_bytes(b"chr1\t1\t50\tg\t2\nchr1\t40\t90\th\t4\n")
    runs = []
    for n in EC_CASES:
        runs.append(("bedmap", ["--ec", "--echo", "--count", "--mean", n]))
        runs.append(("bedmap", ["--ec", "--count", "good.bed", n]))
        runs.append(("bedops", ["--ec", "-m", n]))
        runs.append(("bedops", ["--ec", "-e", "1", n, "good.bed"]))
    runs.append(("bedmap", ["--ec", "--faster", "--count", "nested.bed"]))
    runs.append(("closest-features", ["--ec", "--dist", "space_chrom.bed", "good.bed"]))
    runs.append(("closest-features", ["--ec", "good.bed", "headers_ok.bed"]))
    for tool, argv in runs:
        ours = subprocess.run([bedops_b200.tool_path(tool)] + argv, cwd=tmp_path, capture_output=True)
        ref = subprocess.run([os.path.join(REFBIN, tool)] + argv, cwd=tmp_path, capture_output=True)
        assert ours.returncode == ref.returncode, (tool, argv, ours.stderr, ref.stderr)
        assert ours.stdout == ref.stdout, (tool, argv, ours.stdout, ref.stdout)
        assert ours.stderr == ref.stderr, (tool, argv, ours.stderr, ref.stderr)


def test_closest_features_streaming_state_across_chunks(kit, tmp_path):
    from bedops_b200._lib import COL_LINE
    """Dense nesting on both sides with thousands of reference rows per chromosome: the device runs findDistances
    speculatively in chunks of 256 reference rows (closest.cu) and must reproduce the reference's streaming push-back
    state exactly -- against the oracle's statement-by-statement restatement and, where it travelled, the binary."""
    import numpy as np
    from conftest import have_ref, REFBIN
    import subprocess
    for seed, (nr, nq, span, mu_r, mu_q) in enumerate([(3000, 6000, 40000, 4.0, 3.0), (5000, 20000, 60000, 5.0, 2.5),
                                                       (2500, 2500, 3000, 3.0, 3.0), (4000, 1000, 200000, 6.0, 5.0),
                                                       (6000, 30000, 30000, 2.0, 4.5)]):
        rng = np.random.default_rng(100 + seed)

        def make(n, mu):
            rows = []
            for c in ("chr1", "chr2"):
                s = rng.integers(0, span, n)
                e = s + np.maximum(1, rng.lognormal(mu, 1.2, n).astype(np.int64))
                o = np.lexsort((e, s))
                rows += ["%s\t%d\t%d" % (c, a, b) for a, b in zip(s[o].tolist(), e[o].tolist())]
            return ("\n".join(rows) + "\n").encode()
        rt, qt = make(nr, mu_r), make(nq, mu_q)
        ref, qry = kit.load(rt, 3, COL_LINE), kit.load(qt, 3, COL_LINE)
        for kw in (dict(dist=True), dict(no_overlaps=True, dist=True), dict(closest=True, no_ref=True)):
            got = kit.closest(ref, qry, **kw)
            exp = O.closest_features(rt, qt, **kw)
            if got != exp:
                ge, gg = exp.split(b"\n"), got.split(b"\n")
                bad = [i for i, (x, y) in enumerate(zip(ge, gg)) if x != y]
                raise AssertionError("seed %d %s: %d of %d rows differ, first %d: %r vs %r" % (seed, kw, len(bad), len(ge), bad[0], ge[bad[0]], gg[bad[0]]))
        if have_ref():
            (tmp_path / "r.bed").write_bytes(rt)
            (tmp_path / "q.bed").write_bytes(qt)
            r = subprocess.run([os.path.join(REFBIN, "closest-features"), "--dist", "r.bed", "q.bed"], cwd=tmp_path, capture_output=True)
            assert r.stdout == kit.closest(ref, qry, dist=True)
        ref.free()
        qry.free()


def test_range_shards_with_halos_equal_the_unsharded_call(kit, synth_files, tmp_path):
    """ONE dataset cut at genomic positions inside chromosomes (bk_shard_plan_make), boundary halos found with the
    prefix-max-end index (bk_bed_reach_start / bk_bed_chrom_max_end), halo ++ own records by bk_bed_concat: the parts in
    rank order are the unsharded output, byte for byte.  Ranks run one after the other on this GPU (the exchange is a
    Python list here; bench.py does it with an NCCL allgather, the tools with threads).  Cases: the hg38-shaped pair, a
    single chromosome with an interval spanning every cut, --range padding, list operations that need the map text."""
    from bedops_b200._lib import COL_ID, COL_LINE, COL_SCORE
    from bedops_b200.shard import make_plan, RangeShard
    from test_shard import halo_cases
    for ci, (ref, mp_, ops, overlap) in enumerate(halo_cases(synth_files)):
        ids = any(o.startswith("echo-map") for o in ops)
        fields = 5 if any(o in ("sum", "mean") for o in ops) else (4 if ids else 3)
        mcols = (COL_SCORE if fields == 5 else 0) | ((COL_ID | COL_LINE) if ids else 0)
        rcols = COL_LINE if "echo" in ops else 0
        r0, m0 = kit.load(ref, 3, rcols), kit.load(mp_, fields, mcols)
        whole = kit.bedmap(r0, m0, ops, overlap=overlap)
        r0.free()
        m0.free()
        assert_same(whole, O.bedmap(ref, mp_, ops, overlap=overlap))
        for world in (2, 3, 7):
            plan = make_plan(ref, mp_, world)
            shards = [RangeShard(kit, plan, k, ref, mp_, ops, ref_fields=3, ref_cols=rcols, map_fields=fields, map_cols=mcols,
                                 overlap=overlap) for k in range(world)]
            allr = [s.reach_list() for s in shards]
            parts = [s.finish(allr) for s in shards]
            assert_same(b"".join(parts), whole)
            assert sum(s.bytes_in for s in shards) >= len(ref) + len(mp_)   # every byte is uploaded at least once
    # the tools: BEDKIT_GPUS=N on a single-chromosome input (the whole-chromosome planner could not split it)
    if have_ref():
        import bedops_b200
        import torch
        ref, mp_, ops, overlap = halo_cases(synth_files)[2]
        (tmp_path / "r.bed").write_bytes(ref)
        (tmp_path / "m.bed").write_bytes(mp_)
        env = dict(os.environ, BEDKIT_GPUS="4")
        if torch.cuda.device_count() < 4:
            env["BEDKIT_SHARE_DEVICE"] = "1"
        argv = ["--echo", "--count", "--sum", "--bases", "--echo-map-id", "r.bed", "m.bed"]
        ours = subprocess.run([bedops_b200.tool_path("bedmap")] + argv, cwd=tmp_path, capture_output=True, env=env)
        exp = subprocess.run([os.path.join(REFBIN, "bedmap")] + argv, cwd=tmp_path, capture_output=True)
        assert ours.returncode == 0, ours.stderr
        assert_same(ours.stdout, exp.stdout)
        # a UCSC header under --ec makes "track" a chromosome name that sorts last: the planner must refuse, not duplicate rows
        (tmp_path / "h.bed").write_bytes(b"track name=x\n" + synth_files["r.bed"])
        for tool, argv in (("bedmap", ["--ec", "--count", "h.bed", "h.bed"]), ("bedops", ["--ec", "-m", "h.bed"])):
            ours = subprocess.run([bedops_b200.tool_path(tool)] + argv, cwd=tmp_path, capture_output=True, env=env)
            exp = subprocess.run([os.path.join(REFBIN, tool)] + argv, cwd=tmp_path, capture_output=True)
            assert ours.returncode == exp.returncode, ours.stderr
            assert_same(ours.stdout, exp.stdout)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_row_ids_count_printed_rows(kit, tmp_path):
    """--echo-ref-row-id: the reference's static counter is bumped by every printed id (ProcessBedVisitorRow.hpp:347-354):
    rows dropped by --skip-unmapped or outside --chrom do not count, two id columns take two numbers per row."""
    import bedops_b200
    r = b"chr1\t1\t5\nchr1\t50\t60\nchr1\t100\t110\nchr2\t1\t5\nchr2\t70\t80\nchr3\t5\t9\n"
    m = b"chr1\t2\t3\tx\t1\nchr1\t101\t102\ty\t2\nchr2\t72\t75\tz\t3\n"
    (tmp_path / "r.bed").write_bytes(r)
    (tmp_path / "m.bed").write_bytes(m)
    for argv in (["--echo-ref-row-id", "--count"], ["--skip-unmapped", "--echo-ref-row-id", "--echo"],
                 ["--chrom", "chr2", "--echo-ref-row-id"], ["--skip-unmapped", "--echo-ref-row-id", "--count", "--echo-ref-row-id"],
                 ["--chrom", "chr2", "--skip-unmapped", "--echo-ref-row-id"]):
        full = argv + ["r.bed", "m.bed"]
        exp = subprocess.run([os.path.join(REFBIN, "bedmap")] + full, cwd=tmp_path, capture_output=True)
        got = subprocess.run([bedops_b200.tool_path("bedmap")] + full, cwd=tmp_path, capture_output=True)
        assert got.returncode == 0, got.stderr
        assert_same(got.stdout, exp.stdout)
        assert_same(oracle_cli.run("bedmap", full, {"r.bed": r, "m.bed": m}), exp.stdout)


def test_windows_far_longer_than_a_warp_step(kit):
    """Reference rows whose windows are thousands of map rows long among short ones
    (a row spanning the chromosome; a tile in which every row has a long window): the window scan must not depend on
    the window length.  Every aggregate against the oracle."""
    import numpy as np
    from bedops_b200._lib import COL_LINE, COL_SCORE
    rng = np.random.default_rng(5)
    for n_ref, n_map, span, giant in ((2000, 20000, 400000, 3), (600, 9000, 60000, 600)):
        ms = np.sort(rng.integers(0, span, n_map))
        me = ms + np.maximum(1, rng.lognormal(3.0, 1.0, n_map).astype(np.int64))
        mp_ = "".join("chrQ\t%d\t%d\tm%d\t%d\n" % (a, b, k, k % 1000) for k, (a, b) in enumerate(zip(ms.tolist(), me.tolist()))).encode()
        rs = rng.integers(0, span, n_ref)
        rl = np.maximum(1, rng.lognormal(4.0, 1.0, n_ref).astype(np.int64))
        rl[rng.choice(n_ref, giant, replace=False)] = span   # rows that overlap (nearly) every map row
        order = np.lexsort((rs + rl, rs))
        ref = "".join("chrQ\t%d\t%d\n" % (a, a + l) for a, l in zip(rs[order].tolist(), rl[order].tolist())).encode()
        rb, mb = kit.load(ref, 3, COL_LINE), kit.load(mp_, 5, COL_SCORE)
        for ops in (["echo", "count", "sum", "bases"], ["count", "max", "min", "mean"], ["indicator", "bases"]):
            assert_same(kit.bedmap(rb, mb, ops), O.bedmap(ref, mp_, ops))
        rb.free()
        mb.free()


def test_both_window_kernels_on_chromosome_edges_gaps_and_giant_rows(kit, monkeypatch):
    """The lane-per-row window kernel (k_map_group, the default) and the warp-per-row one (BEDKIT_MAP_KERNEL=row) against
    the oracle where their bookkeeping differs from the common case: many short chromosomes (a batch of 32 reference
    rows and even a group of 8 straddles several), chromosomes missing on either side, windows separated by empty
    stretches (the chunk stream jumps), a reference row longer than 2^26 bases (64 overlaps do not fit 32 bits), a
    reference row covering a whole dense chromosome (finished by the whole warp), every overlap criterion and the list
    operations (which need the exact window)."""
    from bedops_b200._lib import COL_ID, COL_LINE, COL_SCORE
    rng = np.random.default_rng(11)
    ref_rows, map_rows = [], []
    names = ["c%02d" % k for k in range(40)]
    for k, nm in enumerate(names):
        if k % 7 != 3:   # map rows: clusters with gaps of 1e6 between them
            n = int(rng.integers(1, 400))
            base = rng.choice([0, 1_000_000, 5_000_000], n) + rng.integers(0, 3000, n)
            ln = np.maximum(1, rng.lognormal(3.5, 1.2, n).astype(np.int64))
            for a, b in sorted(zip(base.tolist(), (base + ln).tolist())):
                map_rows.append((nm, a, b))
        if k % 5 != 4:   # reference rows: 1..12 per chromosome, some far from any map row
            n = int(rng.integers(1, 13))
            st = np.sort(rng.choice([0, 500_000, 1_000_000, 5_000_000, 9_000_000], n) + rng.integers(0, 3500, n))
            ln = np.maximum(1, rng.lognormal(5.0, 1.5, n).astype(np.int64))
            rr = sorted(zip(st.tolist(), (st + ln).tolist()))
            for a, b in rr:
                ref_rows.append((nm, a, b))
    # a dense chromosome with a reference row over all of it and a giant reference row (> 2^26 bases) among short ones
    dn = 5000
    ds = np.sort(rng.integers(0, 100_000_000, dn))
    de = ds + np.maximum(1, rng.lognormal(8.0, 2.0, dn).astype(np.int64))
    for a, b in sorted(zip(ds.tolist(), de.tolist())):
        map_rows.append(("d00", a, b))
    dref = [(0, 99_000_000), (10, 70_000_000)] + [(int(a), int(a) + 5000) for a in np.sort(rng.integers(0, 100_000_000, 300))]
    for a, b in sorted(dref):
        ref_rows.append(("d00", a, b))
    ref = "".join("%s\t%d\t%d\n" % r for r in ref_rows).encode()
    mp_ = "".join("%s\t%d\t%d\tid%d\t%d\n" % (c, a, b, k, (k * 7919) % 1000 - 300) for k, (c, a, b) in enumerate(map_rows)).encode()
    cases = [(["echo", "count", "sum", "bases"], ("bp", 1)), (["count", "max", "min", "mean"], ("bp", 1)),
             (["count", "bases", "mean"], ("bp", 40)), (["count", "bases", "sum"], ("range", 2000)),
             (["count", "sum"], ("fraction-map", 0.5)), (["count", "min"], ("fraction-ref", 0.1)),
             (["count", "bases"], ("fraction-either", 0.3)), (["count"], ("fraction-both", 0.2)), (["count", "sum"], ("exact", 0)),
             (["echo-map-id", "count"], ("bp", 1)), (["echo-map-id-uniq", "bases"], ("range", 100)),
             (["echo-map-range", "sum"], ("bp", 1))]
    exp = [O.bedmap(ref, mp_, ops, overlap=ov) for ops, ov in cases]
    for env in (None, "row"):
        if env:
            monkeypatch.setenv("BEDKIT_MAP_KERNEL", env)
        else:
            monkeypatch.delenv("BEDKIT_MAP_KERNEL", raising=False)
        rb, mb = kit.load(ref, 3, COL_LINE), kit.load(mp_, 5, COL_SCORE | COL_ID | COL_LINE)
        for (ops, ov), e in zip(cases, exp):
            assert_same(kit.bedmap(rb, mb, ops, overlap=ov), e)
        rb.free()
        mb.free()
