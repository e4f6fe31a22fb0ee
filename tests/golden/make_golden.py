#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference tree and the reference binaries.

Run in the BUILD container only (needs /root/reference and oracle/_ref/bin, see oracle/build_ref.sh):
    python tests/golden/make_golden.py
The GPU box has neither; tests read only the committed JSON.

Outputs
  testplan.json   the reference's own regression plan (applications/bed/bedops/test/TestPlan.xml) replayed in
                  `order` with the reference binary (later tests consume earlier outputs); the 20 in-scope tests
                  (-m -i -e -n) are stored with the bytes of every input file, the argv and the XML's ANSWER.
  docs.json       worked examples transcribed from docs/content/reference/statistics/bedmap.rst and
                  docs/content/reference/set-operations/bedops.rst with their fixture files
                  (docs/assets/reference/statistics/reference_bedmap_{reference,map}.bed) and printed answers.
                  plus the well-formed first 169 rows of reference_bedmap_motifs.bed (the --echo-map-id example's map).
  synthetic.json  sha256 + length of the reference binaries' stdout on seeded synthetic inputs
                  (bedops_b200.synth), so that the oracle stays pinned where the binaries are absent.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import xml.etree.ElementTree as ET

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REF", "/root/reference")
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
sys.path.insert(0, ROOT)

IN_SCOPE = ("-m", "-i", "-e", "-n", "--merge", "--intersect", "--element-of", "--not-element-of",
            "-c", "-d", "-s", "--complement", "--difference", "--symmdiff", "-w", "--chop")


def update_string(s, chrom):
    """Regression.java:183-197: drop spaces, prefix the chromosome."""
    s = s.strip().replace(" ", "")
    out = ""
    for line in s.split("\n"):
        if line == "":
            continue
        out += (chrom + "\t" if chrom else "") + line + "\n"
    return out


def testplan():
    tree = ET.parse(os.path.join(REF, "applications/bed/bedops/test/TestPlan.xml"))
    tests = []
    for t in tree.getroot().findall("TEST"):
        chrom = t.get("chromosome") or ""
        call, inputs, answer, output = "", [], "", None
        for child in t:
            if child.tag == "CALL":
                call += (child.text or "").strip()
            elif child.tag == "OUTPUT":
                output = child.get("name")
            elif child.tag == "INPUT":
                inputs.append((child.get("name"), update_string(child.text or "", chrom)))
            elif child.tag == "ANSWER":
                answer = update_string(child.text or "", chrom) if child.text else ""
        tests.append(dict(order=int(t.get("order")), chrom=chrom, call=call, inputs=inputs, answer=answer, output=output))
    tests.sort(key=lambda d: d["order"])
    kept, passed = [], 0
    with tempfile.TemporaryDirectory() as td:
        for t in tests:
            for name, content in t["inputs"]:
                with open(os.path.join(td, name), "w") as f:
                    f.write(content)
            argv = t["call"].split() + [n for n, _ in t["inputs"]]
            p = subprocess.run([os.path.join(BIN, "bedops"), "--ec"] + argv, cwd=td, capture_output=True, text=True)
            sub = "".join(l.strip() + "\n" for l in p.stdout.split("\n") if len(l) > 0)
            with open(os.path.join(td, t["output"]), "w") as f:
                f.write(sub)
            ok = sub == t["answer"] and p.returncode == 0
            passed += ok
            if not ok:
                print("reference FAILS its own test", t["order"], argv, file=sys.stderr)
            ops = [a for a in argv if a in IN_SCOPE]
            if ops and ok:
                files = {}
                for a in argv:
                    pth = os.path.join(td, a)
                    if os.path.isfile(pth):
                        files[a] = open(pth).read()
                kept.append(dict(order=t["order"], argv=argv, files=files, answer=t["answer"], raw_stdout=p.stdout))
    print("TestPlan.xml: reference passes %d/%d; %d in-scope tests kept" % (passed, len(tests), len(kept)))
    return kept


def docs():
    st = os.path.join(REF, "docs/assets/reference/statistics")
    files = {
        "reference.bed": open(os.path.join(st, "reference_bedmap_reference.bed")).read(),
        "map.bed": open(os.path.join(st, "reference_bedmap_map.bed")).read(),
        "First.bed": "chr1\t100\t200\nchr1\t150\t160\nchr1\t200\t300\nchr1\t400\t475\nchr1\t500\t550\n",
        "Second.bed": "chr1\t120\t125\nchr1\t150\t155\nchr1\t150\t160\nchr1\t460\t470\nchr1\t490\t500\n",
    }
    R = ["chr21\t33031200\t33032400\tref-1", "chr21\t33031400\t33031800\tref-2", "chr21\t33031900\t33032000\tref-3"]

    def rows(vals):
        return "".join(a + "|" + b + "\n" for a, b in zip(R, vals))

    ex = [  # (tool, argv, stdin, expected stdout, source)
        ("bedmap", ["--echo", "--mean", "reference.bed", "map.bed"], None, rows(["43.442623", "31.571429", "154.500000"]), "bedmap.rst:293-298"),
        ("bedmap", ["--mean", "reference.bed", "map.bed"], None, "43.442623\n31.571429\n154.500000\n", "bedmap.rst:310-313"),
        ("bedmap", ["--echo", "--mean", "-", "map.bed"], "chr21\t1000\t2000\tfoo-1\n", "chr21\t1000\t2000\tfoo-1|NAN\n", "bedmap.rst:321-322"),
        ("bedmap", ["--echo", "--count", "--bases", "reference.bed", "map.bed"], None, rows(["61|1200", "21|400", "6|100"]), "bedmap.rst:485-490"),
        ("bedmap", ["--echo", "--indicator", "reference.bed", "map.bed"], None, rows(["1", "1", "1"]), "bedmap.rst:511-516"),
        ("bedmap", ["--echo", "--range", "100", "--mean", "reference.bed", "map.bed"], None, None, "bedmap.rst:584-585 (ref-3 -> 117.750000)"),
        ("bedmap", ["--echo", "--echo-map-id", "-", "motifs.bed"], "chr1\t4534150\t4534300\tref-1\n",
         "chr1\t4534150\t4534300\tref-1|-V_GRE_C;-V_STAT_Q6;+V_HNF4_Q6_01\n", "bedmap.rst:435-436"),
        ("bedops", ["--element-of", "1", "First.bed", "Second.bed"], None, "chr1\t100\t200\nchr1\t150\t160\nchr1\t400\t475\n", "bedops.rst:221-224"),
        ("bedops", ["--element-of", "15", "First.bed", "Second.bed"], None, "chr1\t100\t200\n", "bedops.rst:234-235"),
        ("bedops", ["--element-of", "50%", "First.bed", "Second.bed"], None, "chr1\t150\t160\n", "bedops.rst:247-248"),
    ]
    out = []
    with tempfile.TemporaryDirectory() as td:
        for n, c in files.items():
            open(os.path.join(td, n), "w").write(c)
        # the shipped Motifs fixture has malformed coordinates from line 170 on ("7412.4.5"); keep the well-formed head
        head = open(os.path.join(st, "reference_bedmap_motifs.bed")).read().split("\n")[:169]
        files["motifs.bed"] = "\n".join(head) + "\n"
        open(os.path.join(td, "motifs.bed"), "w").write(files["motifs.bed"])
        for tool, argv, stdin, exp, src in ex:
            p = subprocess.run([os.path.join(BIN, tool)] + argv, cwd=td, input=stdin, capture_output=True, text=True)
            if exp is None:
                exp = p.stdout
                assert "ref-3|117.750000" in exp, exp
            assert p.stdout == exp, (argv, p.stdout, exp)
            out.append(dict(tool=tool, argv=argv, stdin=stdin, expected=exp, source=src))
    print("docs: %d worked examples reproduced by the reference binaries" % len(out))
    return dict(files=files, examples=out)


def synthetic():
    from bedops_b200 import synth
    cases = []
    with tempfile.TemporaryDirectory() as td:
        spec = {"m.bed": (60000, 1, synth.MAP_SHAPE, 5), "m2.bed": (60000, 3, synth.MAP_SHAPE, 5),
                "r.bed": (6000, 2, synth.REF_SHAPE, 5), "m3.bed": (40000, 4, synth.MAP_SHAPE, 3),
                "u.bed": (60000, 1, synth.MAP_SHAPE, 5, True),
                # dense pair on two chromosomes (about 5 hits per reference row) for the per-hit list operations
                "dm.bed": (6000000, 1, synth.MAP_SHAPE, 5, True, ["chr21", "chrY"]),
                "dr.bed": (200000, 2, synth.REF_SHAPE, 5, False, ["chr21", "chrY"])}
        for n, a in spec.items():
            open(os.path.join(td, n), "wb").write(synth.bed_text(a[0], a[1], a[2], a[3], unique=len(a) > 4 and a[4],
                                                                 chroms=a[5] if len(a) > 5 else None))
        runs = [
            ("bedops", ["-m", "m.bed"]), ("bedops", ["-m", "m.bed", "m2.bed", "r.bed", "m3.bed"]),
            ("bedops", ["-i", "r.bed", "m.bed"]), ("bedops", ["-i", "r.bed", "m.bed", "m2.bed"]),
            ("bedops", ["-e", "1", "r.bed", "m.bed"]), ("bedops", ["-e", "m.bed", "r.bed"]),
            ("bedops", ["-e", "50%", "m.bed", "r.bed", "m2.bed"]), ("bedops", ["-n", "25%", "r.bed", "m.bed"]),
            ("bedops", ["-n", "1", "r.bed", "m.bed", "m3.bed"]), ("bedops", ["--chrom", "chr7", "-m", "m.bed", "r.bed"]),
            ("bedmap", ["--echo", "--count", "--mean", "--bases", "r.bed", "m.bed"]),
            ("bedmap", ["--prec", "3", "--delim", "\t", "--count", "--sum", "--max", "--min", "--indicator", "r.bed", "m.bed"]),
            ("bedmap", ["--echo", "--echo-map-id", "r.bed", "u.bed"]),
            ("bedmap", ["--count", "--bases", "m.bed"]),
            ("bedmap", ["--range", "500", "--count", "--bases", "r.bed", "m.bed"]),
            ("bedmap", ["--bp-ovr", "100", "--count", "--mean", "r.bed", "m.bed"]),
            ("bedmap", ["--fraction-ref", "0.5", "--count", "r.bed", "m.bed"]),
            ("bedmap", ["--fraction-map", "0.5", "--count", "r.bed", "m.bed"]),
            ("bedmap", ["--fraction-either", "0.5", "--count", "r.bed", "m.bed"]),
            ("bedmap", ["--fraction-both", "0.5", "--count", "r.bed", "m.bed"]),
            ("bedmap", ["--exact", "--count", "m.bed", "m.bed"]),
            ("bedmap", ["--chrom", "chr2", "--skip-unmapped", "--echo", "--count", "r.bed", "m.bed"]),
            ("bedops", ["-c", "m.bed"]), ("bedops", ["-c", "-L", "m.bed", "r.bed"]), ("bedops", ["-d", "r.bed", "m.bed"]),
            ("bedops", ["-d", "dm.bed", "dr.bed", "m.bed"]), ("bedops", ["-s", "dr.bed", "dm.bed"]),
            ("bedops", ["-s", "m.bed", "m2.bed", "r.bed", "m3.bed"]), ("bedops", ["--chrom", "chr21", "-c", "-L", "dm.bed"]),
            ("bedops", ["-w", "100", "r.bed", "m.bed"]), ("bedops", ["-w", "250", "--stagger", "100", "-x", "r.bed", "m.bed"]),
            ("bedops", ["-w", "100", "--stagger", "30", "r.bed"]), ("bedops", ["--chrom", "chr3", "-w", "r.bed"]),
            ("bedops", ["-u", "m.bed", "m2.bed", "r.bed", "m3.bed"]), ("bedops", ["-u", "dr.bed", "dm.bed"]),
            ("bedops", ["--chrom", "chr7", "-u", "m.bed", "m3.bed"]),
            ("bedmap", ["--echo", "--echo-map", "dr.bed", "dm.bed"]),
            ("bedmap", ["--echo-map", "--mean", "--echo-map-score", "--prec", "2", "dr.bed", "dm.bed"]),
            ("bedmap", ["--echo-map", "--echo-map-id", "--multidelim", ",", "dr.bed", "dm.bed"]),
            ("bedmap", ["--echo-map-size", "--echo-overlap-size", "--echo-map-range", "dr.bed", "dm.bed"]),
            ("bedmap", ["--bases", "--bases-uniq", "--bases-uniq-f", "dr.bed", "dm.bed"]),
            ("bedmap", ["--range", "300", "--count", "--echo-overlap-size", "--bases-uniq", "--bases-uniq-f", "--echo-map-range",
                        "dr.bed", "dm.bed"]),
            ("bedmap", ["--fraction-map", "0.5", "--skip-unmapped", "--echo", "--echo-map", "--bases-uniq", "--delim", "\t",
                        "dr.bed", "dm.bed"]),
            ("bedmap", ["--sci", "--echo-map-score", "--bases-uniq-f", "dr.bed", "dm.bed"]),
            ("bedmap", ["--echo-map-size", "--bases-uniq", "dm.bed"]),
            ("bedmap", ["--echo-map-id-uniq", "--echo-map-id", "--count", "dr.bed", "dm.bed"]),
            ("bedmap", ["--median", "--kth", "0.25", "--kth", "0.9", "--count", "dr.bed", "dm.bed"]),
            ("bedmap", ["--mad", "--mad", "1.4826", "--median", "dr.bed", "dm.bed"]),
            ("bedmap", ["--variance", "--stdev", "--cv", "--mean", "dr.bed", "dm.bed"]),
            ("bedmap", ["--sci", "--prec", "9", "--range", "500", "--stdev", "--cv", "dr.bed", "dm.bed"]),
            ("closest-features", ["--dist", "r.bed", "m.bed"]),
            ("closest-features", ["--closest", "r.bed", "m.bed"]),
            ("closest-features", ["--no-ref", "--dist", "--closest", "r.bed", "m3.bed"]),
        ]
        for tool, argv in runs:
            p = subprocess.run([os.path.join(BIN, tool)] + argv, cwd=td, capture_output=True)
            assert p.returncode == 0, (tool, argv, p.stderr)
            cases.append(dict(tool=tool, argv=argv, sha256=hashlib.sha256(p.stdout).hexdigest(), nbytes=len(p.stdout)))
    print("synthetic: %d reference outputs hashed" % len(cases))
    return dict(files={k: list(v[:2]) + [list(v[2]), v[3]] + list(v[4:]) for k, v in spec.items()}, cases=cases)


if __name__ == "__main__":
    json.dump(testplan(), open(os.path.join(HERE, "testplan.json"), "w"), indent=1)
    json.dump(docs(), open(os.path.join(HERE, "docs.json"), "w"), indent=1)
    json.dump(synthetic(), open(os.path.join(HERE, "synthetic.json"), "w"), indent=1)
