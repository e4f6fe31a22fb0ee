"""bedops --partition and --range L:R: the oracle against the unmodified reference binary (CPU), the
device against the oracle and the drop-in tool against the reference binary (GPU).  Inputs are sort-bed sorted (rest
included) random files with duplicates, nesting and rows near coordinate 0, where --range clamps and re-orders."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REFBIN, have_ref
import bed_oracle as O
import oracle_cli
from test_random_differential import rand_bed

N_CPU = int(os.environ.get("BEDKIT_FUZZ_CPU", "500")) // 2
N_GPU = int(os.environ.get("BEDKIT_FUZZ_GPU", "600")) // 2
OPS = ["-p", "-m", "-i", "-c", "-d", "-s", "-e", "-n", "-u", "-w"]


def resort(text):
    rows = O.parse_bed(text, 3)
    rows.sort(key=lambda r: (r.chrom, r.start, r.end, r.rest3))
    return b"".join(O.echo_b3rest(r) + b"\n" for r in rows)


def make_case(seed):
    rng = np.random.default_rng(seed)
    chroms = [["chr1"], ["chr1", "chr2"], ["chr1", "chr2", "chrX"]][int(rng.integers(0, 3))]
    span = int(rng.choice([30, 300, 5000]))
    n = int(rng.choice([3, 25, 120]))
    op = OPS[int(rng.integers(0, len(OPS)))]
    nf = int(rng.integers(1, 4))
    if op in ("-i", "-d", "-s", "-e", "-n") and nf == 1:
        nf = 2
    lp = int(rng.choice([0, -3, -40, 5, -1, 2, -1000]))
    rp = int(rng.choice([0, 3, -3, 40, -40, 1]))
    if lp < 0 and rp < 0 and -rp > -lp:
        rp = lp          # an end below |R| wraps to 2^64-x in the reference's getFirst(): refused on the device (own test)
    if op == "-u" and lp < 0:
        nf = 1           # clamped ties merge across files in reader-state order: refused on the device (own test)
    files, names = {}, []
    for k in range(nf):
        t = resort(rand_bed(rng, n, span, [c for c in chroms if rng.random() < 0.85] or chroms[:1], fields=int(rng.choice([3, 5]))))
        files["f%d.bed" % k] = t
        names.append("f%d.bed" % k)
    argv = []
    if rng.random() < 0.8 or op == "-p":
        if op != "-p" or rng.random() < 0.5:
            argv += ["--range", "%d:%d" % (lp, rp) if rng.random() < 0.8 else str(abs(lp))]
    if rng.random() < 0.1 and not argv:
        argv += ["--chrom", chroms[0]]
    argv.append(op)
    if op in ("-e", "-n") and rng.random() < 0.7:
        argv.append(str(rng.choice(["1", "5", "50%", "100%"])))
    if op == "-w":
        argv.append(str(int(rng.choice([1, 7, 40]))))
    return argv + names, files


@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_oracle_matches_reference_binary(tmp_path):
    for seed in range(N_CPU):
        argv, files = make_case(20000 + seed)
        for name, data in files.items():
            (tmp_path / name).write_bytes(data)
        r = subprocess.run([os.path.join(REFBIN, "bedops")] + argv, cwd=tmp_path, capture_output=True)
        assert r.returncode == 0, (seed, argv, r.stderr[:300])
        got = oracle_cli.run("bedops", argv, files)
        assert got == r.stdout, (seed, argv, files, got[:300], r.stdout[:300])


def test_padding_quirks_of_the_reference_reader():
    """the documented corner cases of BedPadReader, as golden values taken from the reference binary"""
    rows = O.parse_bed(b"chr1\t2\t10\tb\nchr1\t3\t8\ta\nchr1\t50\t60\tc\nchr2\t1\t30\td\nchr2\t70\t90\te\n", 3)
    # lpad < 0, rpad >= 0: clamped starts, re-ordered by end
    assert O._rows_to_text(O.pad_rows(rows, -5, 0)) == b"chr1\t0\t8\ta\nchr1\t0\t10\tb\nchr1\t45\t60\tc\nchr2\t0\t30\td\nchr2\t65\t90\te\n"
    # lpad < 0 and rpad < 0: clamping only in the zone the constructor read (up to the first survivor beyond |lpad|);
    # chr2's first row wraps below zero and vaporises
    assert O._rows_to_text(O.pad_rows(rows, -5, -2)) == b"chr1\t0\t6\ta\nchr1\t0\t8\tb\nchr1\t45\t58\tc\nchr2\t65\t88\te\n"
    # rows that stop being intervals vaporise
    assert O._rows_to_text(O.pad_rows(rows, 4, -4)) == b"chr1\t54\t56\tc\nchr2\t5\t26\td\nchr2\t74\t86\te\n"


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def kit():
    import bedops_b200
    k = bedops_b200.BedKit(0)
    yield k
    k.close()


@pytest.mark.gpu
def test_device_matches_oracle(kit):
    for seed in range(N_GPU):
        argv, files = make_case(30000 + seed)
        exp = oracle_cli.run("bedops", argv, files)
        got = oracle_cli.run_kit(kit, "bedops", argv, files)
        assert got == exp, (seed, argv, files, got[:300], exp[:300])


@pytest.mark.gpu
def test_device_refuses_what_it_cannot_reproduce(kit):
    from bedops_b200._lib import BedKitError, COL_LINE
    # an end below |R| inside getFirst()'s zone: the reference prints 2^64 - x
    b = kit.load(b"chr1\t5\t9\nchr1\t7\t60\n", 3)
    with pytest.raises(BedKitError) as ei:
        kit.pad(b, -1, -40)
    assert ei.value.code == 5
    # rows of different starts clamped onto equal coordinates, rests out of order, merged across two files
    f1 = kit.load(b"chr1\t1\t9\tz\nchr1\t2\t9\ta\n", 3, COL_LINE)
    f2 = kit.load(b"chr1\t1\t9\tm\n", 3, COL_LINE)
    p1, p2 = kit.pad(f1, -5, 0), kit.pad(f2, -5, 0)
    assert kit.setop("everything", [p1]) == b"chr1\t0\t9\tz\nchr1\t0\t9\ta\n"
    with pytest.raises(BedKitError) as ei:
        kit.setop("everything", [p1, p2])
    assert ei.value.code == 6


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_tool_matches_reference_binary(tmp_path):
    from bedops_b200._lib import tool_path
    for seed in range(40):
        argv, files = make_case(40000 + seed)
        for name, data in files.items():
            (tmp_path / name).write_bytes(data)
        exp = subprocess.run([os.path.join(REFBIN, "bedops")] + argv, cwd=tmp_path, capture_output=True)
        got = subprocess.run([tool_path("bedops")] + argv, cwd=tmp_path, capture_output=True)
        assert (got.returncode, got.stdout) == (exp.returncode, exp.stdout), (seed, argv, got.stderr[:300])
    (tmp_path / "f.bed").write_bytes(b"chr1\t5\t9\n")
    for argv in (["--range", "1:", "-m", "f.bed"], ["--range", "x", "-m", "f.bed"], ["--range", "--3", "-m", "f.bed"],
                 ["--range", "1:2", "--range", "3", "-m", "f.bed"], ["--range"], ["--range", "-2:-2", "-p", "f.bed"]):
        exp = subprocess.run([os.path.join(REFBIN, "bedops")] + argv, cwd=tmp_path, capture_output=True)
        got = subprocess.run([tool_path("bedops")] + argv, cwd=tmp_path, capture_output=True)
        assert (got.returncode, got.stdout, got.stderr) == (exp.returncode, exp.stdout, exp.stderr), argv


@pytest.mark.gpu
def test_partition_at_scale_properties(kit):
    """1 M synthetic rows (BASELINE configs[0] shape): the pieces are disjoint, sorted, cover exactly the merged union, and
    every input coordinate is a piece border."""
    from bedops_b200 import synth
    text = synth.bed_text(1_000_000, 1, synth.MAP_SHAPE)
    b = kit.load(text, 3)
    part = kit.setop("partition", [b])
    merged = kit.setop("merge", [b])
    pb = kit.load(part, 3)
    assert kit.setop("merge", [pb]) == merged                      # same coverage
    assert kit.setop("partition", [pb]) == part                    # idempotent
    s, e, _, _ = pb.columns()
    st, en, _, _ = b.columns()
    import numpy as np
    assert int((e.astype(np.int64) - s).min()) > 0
    # every start of the input opens a piece, every end closes one (per chromosome: compare as sets of (chrom, coord))
    rows = O.parse_bed(text[:2_000_000].rsplit(b"\n", 1)[0] + b"\n", 3)
    first = rows[0].chrom
    sub = [r for r in rows if r.chrom == first]
    pieces = [r for r in O.parse_bed(part[:8_000_000].rsplit(b"\n", 1)[0] + b"\n", 3) if r.chrom == first]
    starts = {r.start for r in pieces}
    ends = {r.end for r in pieces}
    lim = pieces[-1].end
    assert all(r.start in starts for r in sub if r.end <= lim) and all(r.end in ends for r in sub if r.end <= lim)
