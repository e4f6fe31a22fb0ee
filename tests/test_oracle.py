"""CPU tests: pin the oracle (oracle/bed_oracle.py) against the reference's own golden vectors, its worked doc
examples, hashes of the reference binaries' output on seeded synthetic inputs, and -- when oracle/_ref/bin is
present -- live differential runs against the unmodified reference binaries."""
import hashlib
import os
import subprocess

import pytest

from conftest import GOLDEN, REFBIN, have_ref, load_golden
import oracle_cli

TESTPLAN = load_golden("testplan.json")
DOCS = load_golden("docs.json")
SYN = load_golden("synthetic.json")


@pytest.mark.parametrize("case", TESTPLAN, ids=lambda c: "order%d" % c["order"])
def test_oracle_testplan(case):
    files = {k: v.encode() for k, v in case["files"].items()}
    got = oracle_cli.run("bedops", case["argv"], files)
    assert got == case["raw_stdout"].encode()
    # and the XML answer itself, compared the way Regression.java does (trimmed, blank lines dropped)
    norm = "".join(l.strip() + "\n" for l in got.decode().split("\n") if l)
    assert norm == case["answer"]


@pytest.mark.parametrize("ex", DOCS["examples"], ids=lambda e: e["source"].split()[0])
def test_oracle_docs(ex):
    files = {k: v.encode() for k, v in DOCS["files"].items()}
    stdin = ex["stdin"].encode() if ex["stdin"] else None
    assert oracle_cli.run(ex["tool"], ex["argv"], files, stdin) == ex["expected"].encode()


@pytest.mark.parametrize("case", SYN["cases"], ids=lambda c: c["tool"] + "_" + "_".join(c["argv"]).replace("\t", "TAB"))
def test_oracle_synthetic_hashes(case, synth_files):
    got = oracle_cli.run(case["tool"], case["argv"], synth_files)
    assert len(got) == case["nbytes"]
    assert hashlib.sha256(got).hexdigest() == case["sha256"]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_oracle_vs_reference_binaries_live(tmp_path, synth_files):
    for n, c in synth_files.items():
        (tmp_path / n).write_bytes(c)
    for case in SYN["cases"][::4]:
        p = subprocess.run([os.path.join(REFBIN, case["tool"])] + case["argv"], cwd=tmp_path, capture_output=True)
        assert p.returncode == 0
        assert oracle_cli.run(case["tool"], case["argv"], synth_files) == p.stdout
