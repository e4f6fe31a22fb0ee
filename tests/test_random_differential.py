"""Randomised differential tests on small adversarial inputs: duplicates, nesting, touching and abutting intervals,
chromosomes missing from one side, empty files.

  * CPU (-m "not gpu"):  the Python oracle against the unmodified reference binaries (skipped where oracle/_ref/bin is
    absent) -- keeps the oracle pinned on inputs no golden vector covers;
  * GPU (-m gpu):        the CUDA path through the C ABI against the oracle, byte-identical.

Where the reference's output depends on heap addresses (SURVEY 8c hazard 2) the generator avoids the trigger: per-hit
list operations get files without duplicate coordinates.  closest-features is compared on nested / overlapping /
duplicate reference and query rows, with and without --no-overlaps: its streaming state is emulated exactly."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REFBIN, have_ref
import oracle_cli

N_CPU = int(os.environ.get("BEDKIT_FUZZ_CPU", "500"))   # raise for a one-off fuzzing session
N_GPU = int(os.environ.get("BEDKIT_FUZZ_GPU", "600"))


def rand_bed(rng, n, span, chroms, fields=5, unique=False, disjoint=False, messy=False):
    rows = []
    for c in chroms:
        m = int(rng.integers(0, n + 1))
        s = rng.integers(0, span, m)
        l = np.maximum(1, (rng.lognormal(2.0, 1.2, m)).astype(np.int64))
        e = s + l
        order = np.lexsort((e, s))
        s, e = s[order], e[order]
        if disjoint and m:
            keep, last = [], -1
            for a, b in zip(s.tolist(), e.tolist()):
                if a >= last:
                    keep.append((a, b))
                    last = b
            s, e = np.array([k[0] for k in keep], dtype=np.int64), np.array([k[1] for k in keep], dtype=np.int64)
        if unique and len(s):
            k = np.ones(len(s), dtype=bool)
            k[1:] = (s[1:] != s[:-1]) | (e[1:] != e[:-1])
            s, e = s[k], e[k]
        for a, b in zip(s.tolist(), e.tolist()):
            sa, sb, sep1, sep2 = "%d" % a, "%d" % b, "\t", "\t"
            if messy and rng.random() < 0.25:   # what fscanf("%s\t%lu\t%lu...") accepts besides the canonical form
                v = int(rng.integers(0, 5))
                if v == 0:
                    sa = "000" + sa
                elif v == 1:
                    sb = "+" + sb
                elif v == 2:
                    sep1 = "  "
                elif v == 3:
                    sep2 = " \t "
                else:
                    sa, sb = "+0" + sa, "0" + sb
            if fields == 3:
                rows.append("%s%s%s%s%s" % (c, sep1, sa, sep2, sb))
            else:
                rows.append("%s%s%s%s%s\tid%d\t%d" % (c, sep1, sa, sep2, sb, int(rng.integers(0, 50)), int(rng.integers(0, 100))))
            if messy and rng.random() < 0.03:
                rows.append("")                 # blank lines are not records
    if not rows:
        return b""
    eol = "\r\n" if messy and rng.random() < 0.15 else "\n"   # DOS line ends: the '\r' stays in the rest of the line
    text = eol.join(rows) + eol
    if messy and rng.random() < 0.1:
        text = text[:-1]                       # an unterminated last line is not a record either
    return text.encode()


BEDMAP_SCORE = ["--sum", "--mean", "--max", "--min", "--variance", "--stdev", "--cv", "--median", "--kth 0.3", "--kth 0.75", "--mad", "--mad 2.5"]
BEDMAP_PLAIN = ["--echo", "--count", "--indicator", "--bases", "--echo-ref-size", "--echo-ref-name", "--bases-uniq",
                "--bases-uniq-f", "--echo-map-size", "--echo-overlap-size", "--echo-map-range"]
BEDMAP_LIST = ["--echo-map", "--echo-map-id", "--echo-map-score", "--echo-map-id-uniq"]
OVERLAPS = [[], ["--bp-ovr", "3"], ["--range", "5"], ["--fraction-ref", "0.5"], ["--fraction-map", "0.4"],
            ["--fraction-either", "0.6"], ["--fraction-both", "0.3"], ["--exact"]]


def make_case(seed, for_binary):
    rng = np.random.default_rng(seed)
    chroms = [["chr1"], ["chr1", "chr2"], ["chr1", "chr2", "chrX"]][int(rng.integers(0, 3))]
    sub = lambda: [c for c in chroms if rng.random() < 0.85] or chroms[:1]
    span = int(rng.choice([60, 300, 5000]))
    n = int(rng.choice([3, 25, 120]))
    kind = int(rng.integers(0, 10))
    messy = bool(rng.random() < 0.3)
    files = {}
    if kind <= 4:  # bedmap
        ops = list(rng.choice(BEDMAP_PLAIN, int(rng.integers(1, 4)), replace=False))
        ops += list(rng.choice(BEDMAP_SCORE, int(rng.integers(0, 3)), replace=False))
        lists = list(rng.choice(BEDMAP_LIST, int(rng.integers(0, 2)), replace=False))
        ops += lists
        rng.shuffle(ops)
        ops = [t for o in ops for t in o.split(" ")]   # "--kth 0.3" -> two arguments
        argv = list(OVERLAPS[int(rng.integers(0, len(OVERLAPS)))])
        if rng.random() < 0.3:
            argv += ["--prec", str(int(rng.integers(0, 9)))]
        if rng.random() < 0.2:
            argv += ["--sci"]
        if rng.random() < 0.2:
            argv += ["--skip-unmapped"]
        if rng.random() < 0.2:
            argv += ["--delim", "\t"]
        if rng.random() < 0.2:
            argv += ["--multidelim", ","]
        files["r.bed"] = rand_bed(rng, n, span, sub(), messy=messy)
        files["m.bed"] = rand_bed(rng, 2 * n, span, sub(), unique=bool(lists), messy=messy)
        if rng.random() < 0.15 and not lists:   # single-file mode
            return "bedmap", argv + ops + ["m.bed"], files
        return "bedmap", argv + ops + ["r.bed", "m.bed"], files
    if kind <= 8:  # bedops
        nf = int(rng.integers(1, 4))
        op = ["-m", "-c", "-u", "-w"][int(rng.integers(0, 4))] if nf == 1 else \
            ["-m", "-i", "-e", "-n", "-c", "-d", "-s", "-u", "-w"][int(rng.integers(0, 9))]
        names = []
        for k in range(nf):  # -u re-sorts its rows below (line = row): keep those files canonical
            files["f%d.bed" % k] = rand_bed(rng, n, span, sub(), fields=int(rng.choice([3, 5])), messy=messy and op != "-u")
            names.append("f%d.bed" % k)
        argv = [op]
        if op in ("-e", "-n") and rng.random() < 0.7:
            argv.append(str(rng.choice(["1", "5", "50%", "100%", "10%"])))
        if op == "-c" and rng.random() < 0.5:
            argv.append("-L")
        if op == "-w":
            argv.append(str(int(rng.choice([1, 7, 40]))))
            if rng.random() < 0.5:
                argv += ["--stagger", str(int(rng.choice([1, 3, 25])))]
            if rng.random() < 0.4:
                argv.append("-x")
        if op == "-u":  # sort-bed order includes the rest of the line: make the inputs honour it
            for k in names:
                import bed_oracle as O
                rows = O.parse_bed(files[k], 3)
                lines = files[k].split(b"\n")[:-1]
                key = sorted(range(len(rows)), key=lambda i: (rows[i].chrom, rows[i].start, rows[i].end, rows[i].rest3))
                files[k] = b"".join(lines[i] + b"\n" for i in key)
        if rng.random() < 0.15:
            argv = ["--chrom", chroms[0]] + argv
        return "bedops", argv + names, files
    # closest-features
    # nested / overlapping / duplicate rows on both sides: the streaming push-back state of findDistances is part of
    # the contract (oracle: closest_pairs)
    files["r.bed"] = rand_bed(rng, n, span, sub(), messy=messy)
    files["q.bed"] = rand_bed(rng, 2 * n, span, sub(), messy=messy)
    argv = [a for a in ("--dist", "--closest", "--no-ref") if rng.random() < 0.5]
    if rng.random() < 0.35:
        argv.append("--no-overlaps")
    if rng.random() < 0.1:
        argv += ["--chrom", chroms[0]]
    return "closest-features", argv + ["r.bed", "q.bed"], files


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_oracle_matches_the_reference_binaries_on_random_inputs(tmp_path):
    for seed in range(N_CPU):
        tool, argv, files = make_case(seed, for_binary=True)
        if tool == "closest-features" and any(len(v) == 0 for v in files.values()):
            continue
        for name, data in files.items():
            (tmp_path / name).write_bytes(data)
        r = subprocess.run([os.path.join(REFBIN, tool)] + argv, cwd=tmp_path, capture_output=True)
        assert r.returncode == 0, (seed, tool, argv, r.stderr[:300])
        got = oracle_cli.run(tool, argv, files)
        assert got == r.stdout, (seed, tool, argv, files, got[:300], r.stdout[:300])


@pytest.mark.gpu
def test_device_matches_the_oracle_on_random_inputs():
    import bedops_b200
    kit = bedops_b200.BedKit(0)
    try:
        for seed in range(1000, 1000 + N_GPU):
            tool, argv, files = make_case(seed, for_binary=False)
            exp = oracle_cli.run(tool, argv, files)
            got = oracle_cli.run_kit(kit, tool, argv, files)
            assert got == exp, (seed, tool, argv, files, got[:300], exp[:300])
    finally:
        kit.close()


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_command_lines_match_the_reference_binaries_on_random_inputs(tmp_path):
    """argv -> stdout bytes and exit code, our tools against the reference's, on the same random cases (covers the
    option grammar of every operation the generator knows, including --kth <val>, -w [bp] --stagger nt -x, -c -L)."""
    from bedops_b200._lib import tool_path
    for seed in range(5000, 5040):
        tool, argv, files = make_case(seed, for_binary=True)
        if tool == "closest-features" and any(len(v) == 0 for v in files.values()):
            continue
        for name, data in files.items():
            (tmp_path / name).write_bytes(data)
        exp = subprocess.run([os.path.join(REFBIN, tool)] + argv, cwd=tmp_path, capture_output=True)
        got = subprocess.run([tool_path(tool)] + argv, cwd=tmp_path, capture_output=True)
        assert got.returncode == exp.returncode, (seed, tool, argv, got.stderr[:300], exp.stderr[:300])
        assert got.stdout == exp.stdout, (seed, tool, argv, files, got.stdout[:300], exp.stdout[:300])
