"""bedmap --wmean, --tmean <low> <hi>, --kth 0|1, --max-element, --min-element: the oracle against the
unmodified reference binary (CPU), the device against the oracle and the drop-in tool against the reference (GPU).

--wmean adds in heap-address order in the reference (std::set<MapType*>): the comparison with the binary allows the last
printed digit to differ.  --tmean with 0 < low and round(low*n) == 0 reads a marker that earlier rows left behind
(DESIGN.md parity notes): the generator keeps to low == 0 and to low + hi == 1 (the k-th element form)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REFBIN, have_ref
import bed_oracle as O
import oracle_cli
from test_random_differential import rand_bed, OVERLAPS

N_CPU = int(os.environ.get("BEDKIT_FUZZ_CPU", "500")) // 2
N_GPU = int(os.environ.get("BEDKIT_FUZZ_GPU", "600")) // 2
NEW_OPS = ["--wmean", "--tmean 0 0.2", "--tmean 0 0.5", "--tmean 0 0", "--tmean 0.5 0.5", "--tmean 0.3 0.7", "--kth 0", "--kth 1",
           "--max-element", "--min-element"]
OLD_OPS = ["--echo", "--count", "--mean", "--echo-map-id"]


def make_case(seed):
    rng = np.random.default_rng(seed)
    chroms = [["chr1"], ["chr1", "chr2"], ["chr1", "chr2", "chrX"]][int(rng.integers(0, 3))]
    span = int(rng.choice([60, 300, 5000]))
    n = int(rng.choice([3, 25, 120]))
    ops = list(rng.choice(NEW_OPS, int(rng.integers(1, 4)), replace=False)) + list(rng.choice(OLD_OPS, int(rng.integers(0, 3)), replace=False))
    rng.shuffle(ops)
    element = any(o.endswith("-element") for o in ops)
    argv = list(OVERLAPS[int(rng.integers(0, len(OVERLAPS)))])
    if rng.random() < 0.3:
        argv += ["--prec", str(int(rng.integers(0, 9)))]
    if rng.random() < 0.15:
        argv += ["--sci"]
    if rng.random() < (0.7 if element else 0.2):
        argv += ["--skip-unmapped"]
    if rng.random() < 0.2:
        argv += ["--delim", "\t"]
    files = {"r.bed": rand_bed(rng, n, span, chroms[:1] + [c for c in chroms[1:] if rng.random() < 0.85]),
             "m.bed": rand_bed(rng, 2 * n, span, [c for c in chroms if rng.random() < 0.85] or chroms[:1], unique=True)}
    return argv + [t for o in ops for t in o.split(" ")] + ["r.bed", "m.bed"], files


def oracle_run(argv, files):
    try:
        return 0, oracle_cli.run("bedmap", argv, files), b""
    except O.NanElement as e:
        return 1, e.stdout, b"May use bedmap --help for more help.\n\nError: Unable to process a 'NAN' with PrintAllScorePrecision.\n"


def close_enough(a: bytes, b: bytes) -> bool:
    """equal, or equal up to the last printed digit of a floating-point column (--wmean's summation order)"""
    if a == b:
        return True
    la, lb = a.split(b"\n"), b.split(b"\n")
    if len(la) != len(lb):
        return False
    import re
    for x, y in zip(la, lb):
        if x == y:
            continue
        fx, fy = re.split(rb"[|\t;]", x), re.split(rb"[|\t;]", y)
        if len(fx) != len(fy):
            return False
        for p, q in zip(fx, fy):
            if p == q:
                continue
            try:
                u, v = float(p), float(q)
            except ValueError:
                return False
            if abs(u - v) > 1e-9 * max(1.0, abs(u), abs(v)) + 1.0000001 * 10.0 ** -len(p.split(b".")[-1].split(b"e")[0]) * (1 if b"e" not in p else abs(u)):
                return False
    return True


@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_oracle_matches_reference_binary(tmp_path):
    for seed in range(N_CPU):
        argv, files = make_case(50000 + seed)
        for name, data in files.items():
            (tmp_path / name).write_bytes(data)
        r = subprocess.run([os.path.join(REFBIN, "bedmap")] + argv, cwd=tmp_path, capture_output=True)
        rc, out, err = oracle_run(argv, files)
        assert (r.returncode, r.stderr) == (rc, err), (seed, argv, r.stderr[:200])
        assert close_enough(out, r.stdout) if "--wmean" in argv else out == r.stdout, (seed, argv, files, out[:300], r.stdout[:300])


def test_extreme_element_tie_rules():
    """golden values from the reference binary: equal scores go to the genomically last row for --max-element and to the
    first for --min-element; an unmapped row ends the run after the delimiter"""
    ref = b"chr1\t10\t20\tr1\nchr1\t100\t200\tr2\nchr1\t150\t160\tr3\n"
    mp = b"chr1\t5\t15\ta\t3\tx y\nchr1\t12\t18\tb\t7\nchr1\t12\t18\tc\t7\tzz\nchr1\t14\t30\td\t7\nchr1\t120\t130\te\t1\n"
    with pytest.raises(O.NanElement) as ei:
        O.bedmap(ref, mp, ["echo", "max-element", "min-element"])
    assert ei.value.stdout == (b"chr1\t10\t20\tr1|chr1\t14\t30\td\t7.000000|chr1\t5\t15\ta\t3.000000\tx y\n"
                               b"chr1\t100\t200\tr2|chr1\t120\t130\te\t1.000000|chr1\t120\t130\te\t1.000000\n"
                               b"chr1\t150\t160\tr3|")
    assert O.bedmap(ref, mp, ["wmean", "tmean:0.2:0.8", "max"]) == b"6.130435|3.000000|7.000000\n1.000000|1.000000|1.000000\nNAN|NAN|NAN\n"


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_matches_oracle():
    import bedops_b200
    from bedops_b200._lib import BedKitError
    kit = bedops_b200.BedKit(0)
    try:
        for seed in range(N_GPU):
            argv, files = make_case(60000 + seed)
            rc, exp, _ = oracle_run(argv, files)
            if rc == 0:
                got = oracle_cli.run_kit(kit, "bedmap", argv, files)
                assert got == exp, (seed, argv, files, got[:300], exp[:300])
            else:
                with pytest.raises(BedKitError) as ei:
                    oracle_cli.run_kit(kit, "bedmap", argv, files)
                assert ei.value.code == 10, (seed, argv)
    finally:
        kit.close()


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_tool_matches_reference_binary(tmp_path):
    from bedops_b200._lib import tool_path
    for seed in range(60):
        argv, files = make_case(70000 + seed)
        for name, data in files.items():
            (tmp_path / name).write_bytes(data)
        exp = subprocess.run([os.path.join(REFBIN, "bedmap")] + argv, cwd=tmp_path, capture_output=True)
        got = subprocess.run([tool_path("bedmap")] + argv, cwd=tmp_path, capture_output=True)
        assert (got.returncode, got.stderr) == (exp.returncode, exp.stderr), (seed, argv, got.stderr[:300])
        assert close_enough(got.stdout, exp.stdout) if "--wmean" in argv else got.stdout == exp.stdout, (seed, argv, got.stdout[:300], exp.stdout[:300])
    (tmp_path / "r.bed").write_bytes(b"chr1\t5\t9\n")
    (tmp_path / "m.bed").write_bytes(b"chr1\t5\t9\ta\t1\n")
    for argv in (["--tmean", "0.5", "0.6", "r.bed", "m.bed"], ["--tmean", "0.5", "r.bed", "m.bed"], ["--tmean", "x", "0.1", "r.bed", "m.bed"],
                 ["--tmean", "2", "0.1", "r.bed", "m.bed"], ["--kth", "1.5", "r.bed", "m.bed"], ["--kth", "1", "r.bed", "m.bed"],
                 ["--max-element", "m.bed"], ["--min-element", "--echo", "m.bed"]):
        exp = subprocess.run([os.path.join(REFBIN, "bedmap")] + argv, cwd=tmp_path, capture_output=True)
        got = subprocess.run([tool_path("bedmap")] + argv, cwd=tmp_path, capture_output=True)
        assert (got.returncode, got.stdout, got.stderr) == (exp.returncode, exp.stdout, exp.stderr), argv
