"""The argv grammar of the four drop-in tools against the reference binaries, without a GPU (the process
boundary -- argv in, stderr text and exit code out -- is the drop-in contract).

Random command lines (seeded) made of real options, near-misses, stray words and values of every shape are given to our
tool and to the unmodified reference tool in an empty directory: whatever a command line is refused for -- or whichever
file it then fails to find -- both must say the same thing with the same exit code.  Command lines that get as far as the
device are only comparable on a GPU box (tests/test_random_differential.py does that); here they are skipped.

This is how the reference's quirks were found that the tools now reproduce: bedmap builds the "Apparent option" complaint
from the word AFTER the value (and dies on a null pointer when there is none, Input.hpp:116-133), reads `--mad`'s optional
multiplier without looking whether argv has ended (:275), converts numbers with `stream >> member` (an empty word leaves
the default), and says "Unable to find: X" instead of "Unable to find file: X" under --ec (Bedmap.cpp:231, :300-302).
"""
import os
import random
import re
import subprocess

import pytest

from conftest import REFBIN, have_ref

VALUES = ["0", "1", "5", "100", "0.5", "1.0", "0.0", "-1", "1.5", "abc", "--x", "\\t", "|", ";", "chr1", "all", "", "-", "0.25", ".5",
          "1e3", "99999999999", "-0.5", "0.9", "7", "50%", "100%", "0%", "101%", "10:20", "-5:5", "5:-5", "a:b", "1:", "12%x", "%"]
FILES = ["a.bed", "b.bed", "-", "nofile", "c.bed"]

BEDMAP_OPS = ["bases", "bases-uniq", "bases-uniq-f", "echo", "echo-ref-size", "echo-ref-name", "echo-ref-row-id", "echo-map",
              "echo-map-id", "echo-map-id-uniq", "echo-map-size", "echo-overlap-size", "echo-map-range", "echo-map-score", "count",
              "indicator", "max", "max-element", "min", "min-element", "mean", "variance", "stdev", "cv", "sum", "wmean", "median",
              "mad", "kth", "tmean"]
BEDMAP_FLAGS = ["ec", "header", "faster", "sweep-all", "skip-unmapped", "sci", "exact"]
BEDMAP_VALUED = ["delim", "multidelim", "chrom", "prec", "bp-ovr", "range", "fraction-ref", "fraction-map", "fraction-either",
                 "fraction-both"]


def gen_bedmap(rng):
    a = []
    for _ in range(rng.randrange(0, 7)):
        r = rng.random()
        if r < 0.45:
            o = rng.choice(BEDMAP_OPS)
            a.append("--" + o)
            if o in ("mad", "kth", "tmean") and rng.random() < 0.8:
                a += [rng.choice(VALUES) for _ in range(rng.randrange(0, 3))]
        elif r < 0.6:
            a.append("--" + rng.choice(BEDMAP_FLAGS))
        elif r < 0.92:
            a.append("--" + rng.choice(BEDMAP_VALUED))
            if rng.random() < 0.9:
                a.append(rng.choice(VALUES))
        elif r < 0.96:
            a.append(rng.choice(["--bogus", "-x", "x--y", "plain", "--"]))
        else:
            a.append(rng.choice(VALUES))
    return a + [rng.choice(FILES) for _ in range(rng.choice([0, 1, 2, 2, 2, 3]))]


def gen_bedops(rng):
    ops = ["-m", "--merge", "-i", "--intersect", "-e", "--element-of", "-n", "--not-element-of", "-c", "--complement", "-d",
           "--difference", "-s", "--symmdiff", "-p", "--partition", "-u", "--everything", "-w", "--chop", "-L", "--stagger", "-x",
           "--ec", "--header", "--chrom", "--range", "--help-merge", "-h", "--bogus", "-z", "--exclude"]
    a = []
    for _ in range(rng.randrange(0, 5)):
        a.append(rng.choice(ops))
        if rng.random() < 0.5:
            a.append(rng.choice(VALUES))
    return a + [rng.choice(FILES) for _ in range(rng.choice([0, 1, 2, 2, 3]))]


def gen_closest(rng):
    ops = ["--dist", "--closest", "--no-ref", "--no-overlaps", "--ec", "--header", "--chrom", "--delim", "--center", "--shortest",
           "--bogus", "-x", "--no-query", "--print-dist"]
    a = []
    for _ in range(rng.randrange(0, 5)):
        a.append(rng.choice(ops))
        if rng.random() < 0.3:
            a.append(rng.choice(VALUES))
    return a + [rng.choice(FILES) for _ in range(rng.choice([0, 1, 2, 2, 2, 3]))]


def gen_sort(rng):
    ops = ["--max-mem", "--tmpdir", "--check-sort", "--bogus", "-x", "--unique", "--duplicates"]
    a = []
    for _ in range(rng.randrange(0, 4)):
        a.append(rng.choice(ops))
        if rng.random() < 0.5:
            a.append(rng.choice(VALUES + ["2G", "500M", "3K", "1T", "G", "/tmp", "/nonexistent"]))
    return a + [rng.choice(FILES) for _ in range(rng.choice([0, 1, 1, 2, 3]))]


def run(binary, argv, cwd, env=None):
    p = subprocess.run([binary] + argv, cwd=cwd, capture_output=True, stdin=subprocess.DEVNULL, timeout=60, env=env)
    return p.returncode, p.stdout, p.stderr


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
@pytest.mark.parametrize("tool,gen,cases", [("bedmap", gen_bedmap, 600), ("bedops", gen_bedops, 500),
                                            ("closest-features", gen_closest, 300), ("sort-bed", gen_sort, 300)])
def test_refused_command_lines_are_refused_in_the_reference_s_words(tmp_path, tool, gen, cases):
    import bedops_b200
    ours, ref = bedops_b200.tool_path(tool), os.path.join(REFBIN, tool)
    rng = random.Random(20260 + len(tool))
    compared = 0
    for _ in range(cases):
        argv = gen(rng)
        got = run(ours, argv, tmp_path)
        if b"CUDA failure" in got[2] or b"no usable sm_100" in got[2]:
            # a complete command line over readable input (here: stdin alone) is the device's business -- but then the
            # reference must not have refused it either (sort-bed --check-sort reads stdin before it misses a later file)
            if not (tool == "sort-bed" and "--check-sort" in argv):
                assert run(ref, argv, tmp_path)[0] == 0, (tool, argv)
            continue
        if tool == "bedmap" and any("-element" in w for w in argv) and got[0] == 0:
            continue
        exp = run(ref, argv, tmp_path)
        compared += 1
        assert got == exp, (tool, argv)
    assert compared > cases // 2


def test_bedmap_command_lines_are_understood_as_the_reference_understands_them(tmp_path):
    """BEDKIT_DUMP_OPTIONS=1 prints what parse_args made of argv and stops before the library is touched."""
    import bedops_b200
    env = dict(os.environ, BEDKIT_DUMP_OPTIONS="1")
    tool = bedops_b200.tool_path("bedmap")

    def dump(*argv):
        rc, out, err = run(tool, list(argv), tmp_path, env)
        assert rc == 0, err
        return dict(kv.split("=", 1) for kv in out.decode().strip().split(" "))

    d = dump("--echo", "--count", "--mean", "--bases", "r.bed", "m.bed")
    assert (d["ref"], d["map"], d["files"], d["overlap"], d["overlap_bp"], d["prec"], d["map_fields"]) == ("r.bed", "m.bed", "2", "0", "1", "6", "5")
    assert d["ops"] == "1(0,0),2(0,0),6(0,0),4(0,0)"
    # --range 0 is --bp-ovr 1; --kth 0 / 1 are --min / --max; --mad takes a multiplier only when one follows
    d = dump("--range", "0", "--kth", "1", "--kth", "0", "--mad", "--mad", "2.5", "--tmean", "0.1", "0.2", "m.bed")
    assert (d["files"], d["overlap"], d["overlap_bp"], d["ref_fields"], d["map_fields"]) == ("1", "0", "1", "5", "0")
    ops = lambda dumped: re.findall(r"\d+\([^)]*\)", dumped["ops"])
    mx, mn, mad = ops(dump("--max", "--min", "--mad", "x"))
    number = lambda op: op.split("(")[0]
    assert [number(x) for x in ops(d)[:3]] == [number(mx), number(mn), number(mad)]
    assert ops(d)[3].endswith("(2.5,0)") and ops(d)[4].endswith("(0.10000000000000001,0.20000000000000001)")
    # an empty word converts to nothing: the default stays (stream >> member, Input.hpp:139-141)
    assert dump("--prec", "", "--sum", "r.bed", "m.bed")["prec"] == "6"
    assert dump("--prec", "3", "--sci", "--fraction-both", "0.5", "--faster", "--sum", "r.bed", "m.bed")["frac"] == "0.5"
    assert dump("--delim", "\\t", "--multidelim", ",", "--chrom", "chrX", "--skip-unmapped", "--count", "-", "m.bed")["chrom"] == "[chrX]"
