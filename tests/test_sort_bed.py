"""sort-bed: oracle vs the unmodified reference binary (CPU), device vs oracle and the drop-in tool vs
the reference binary (GPU), on shuffled rows with duplicate coordinates, rows that differ only in their rest, space
separators, leading zeros, header zones, empty lines, several files -- and on every line the reference rejects."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REFBIN, have_ref
import bed_oracle as O

N_CPU = int(os.environ.get("BEDKIT_FUZZ_CPU", "500")) // 5
N_GPU = int(os.environ.get("BEDKIT_FUZZ_GPU", "600")) // 4


def rand_unsorted(rng, n, messy):
    chroms = ["chr1", "chr10", "chr2", "chrX", "scaffold_%d" % int(rng.integers(0, 1000)), "c" * int(rng.integers(1, 40))]
    chroms = chroms[:int(rng.integers(1, len(chroms) + 1))]
    rows = []
    span = int(rng.choice([20, 1000, 250_000_000]))
    for _ in range(n):
        c = chroms[int(rng.integers(0, len(chroms)))]
        s = int(rng.integers(0, span))
        e = s + int(rng.integers(1, 6 if rng.random() < 0.5 else 5000))
        sa, sb, s1, s2 = "%d" % s, "%d" % e, "\t", "\t"
        kind = int(rng.integers(0, 6))
        rest = ""
        if kind == 1:
            rest = "\tid%d" % int(rng.integers(0, 4))
        elif kind == 2:
            rest = "\tid%d\t%d" % (int(rng.integers(0, 3)), int(rng.integers(0, 3)))
        elif kind == 3:
            rest = "\t" + "".join(rng.choice(list("ab \t.+"), int(rng.integers(1, 6))))
        elif kind == 4:
            rest = " x%d" % int(rng.integers(0, 3))       # a space separates the rest as well
        if messy and rng.random() < 0.3:
            v = int(rng.integers(0, 4))
            if v == 0:
                sa = "00" + sa
            elif v == 1:
                s1 = " "
            elif v == 2:
                s2 = " "
            else:
                rest = rest + "\t" if rest else "\t \t"   # trailing white space: no rest
        rows.append(c + s1 + sa + s2 + sb + rest)
        if messy and rng.random() < 0.05:
            rows.append("")
    head = []
    if messy and rng.random() < 0.4:
        head = [["track name=x", "browser position chr1", "#comment", "@HD\tVN", ""][int(rng.integers(0, 5))]
                for _ in range(int(rng.integers(1, 4)))]
    text = "\n".join(head + rows) + "\n"
    if messy and rows and rng.random() < 0.15 and ("\t" in rows[-1][rows[-1].find("\t") + 1:] and rows[-1].count("\t") >= 3):
        text = text[:-1]                                   # unterminated last line with a rest: still a row
    return text.encode()


BAD_LINES = [
    b" chr1\t1\t2\n", b"\tchr1\t1\t2\n", b"chr1\n", b"chr1\t5\n", b"chr1\t\t5\t6\n", b"chr1\t5a\t6\n", b"chr1\t+5\t6\n",
    b"chr1\t5\t\t6\n", b"chr1\t5\t6x\n", b"chr1\t5\t6\r\n", b"chr1\t6\t6\n", b"chr1\t7\t6\n", b"chr1\t1234567890123\t5\n",
    b"chr1\t5\t1234567890123\n", b"c" * 128 + b"\t1\t2\n", b"chr1\t1\t2\t" + b"i" * 16384 + b"\n", b"chr1  5 6\n",
    b"   \n", b"chr1\t1\t2", b"track\t1\n",
]


def run_ref(tool, argv, files, tmp_path, bindir=REFBIN):
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
    p = subprocess.run([os.path.join(bindir, tool)] + argv, cwd=tmp_path, capture_output=True)
    return p.returncode, p.stdout, p.stderr


def oracle_run(files, names):
    try:
        return 0, O.sort_bed([files[n] for n in names], names), b""
    except O.SortBedError as e:
        return 1, b"", e.message.encode()


@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_oracle_matches_reference_binary(tmp_path):
    for seed in range(N_CPU):
        rng = np.random.default_rng(7000 + seed)
        nf = int(rng.integers(1, 4))
        files = {"f%d.bed" % k: rand_unsorted(rng, int(rng.choice([0, 3, 40, 300])), bool(rng.random() < 0.5)) for k in range(nf)}
        names = sorted(files)
        rc, out, err = run_ref("sort-bed", names, files, tmp_path)
        assert (rc, out, err) == oracle_run(files, names), "seed %d" % seed


@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_oracle_error_messages_match_reference_binary(tmp_path):
    good = b"chr2\t5\t9\tz\nchr1\t1\t2\n"
    for k, bad in enumerate(BAD_LINES):
        for files in ({"a.bed": good + bad}, {"a.bed": good, "b.bed": b"#h\n\n" + bad + (good if bad.endswith(b"\n") else b"")}):
            names = sorted(files)
            rc, out, err = run_ref("sort-bed", names, files, tmp_path)
            assert (rc, out, err) == oracle_run(files, names), "bad line %d %r" % (k, bad[:40])
            assert rc == 1 or (bad.startswith(b"track") and len(files) == 2)   # a header line in the header zone is skipped


@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_tool_argv_handling_without_a_device(tmp_path):
    """everything the tool answers before it needs the GPU: banners, usage, --max-mem grammar, missing files"""
    from bedops_b200._lib import tool_path
    ours = os.path.dirname(tool_path("sort-bed"))
    for argv in ([], ["--help"], ["--version"], ["--max-mem"], ["--max-mem", "10", "f.bed"], ["--max-mem", "1T", "f.bed"],
                 ["--max-mem", "G", "f.bed"], ["--max-mem", "1G", "--max-mem", "2G", "f.bed"], ["--tmpdir"],
                 ["--tmpdir", "a", "--tmpdir", "b", "f.bed"], ["--max-mem", "1G"], ["--check-sort"], ["-", "-"],
                 ["nonexistent.bed"], ["--help", "nonexistent.bed"]):
        assert run_ref("sort-bed", argv, {}, tmp_path, ours) == run_ref("sort-bed", argv, {}, tmp_path), argv


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def kit():
    import bedops_b200
    k = bedops_b200.BedKit(0)
    yield k
    k.close()


def strip_headers(text):
    """what the tool does before the text goes to the library: the header zone at the top of a file is dropped"""
    lines = text.split(b"\n")
    k = 0
    while k < len(lines) - 1:
        ln = lines[k]
        if ln == b"" or (ln[:1] not in (b" ", b"\t") and (ln.startswith(b"browser") or ln.startswith(b"track") or ln[:1] in (b"#", b"@"))):
            k += 1
            continue
        break
    body = b"\n".join(lines[k:])
    return body if body.endswith(b"\n") or not body else body + b"\n"


@pytest.mark.gpu
def test_device_sort_matches_oracle(kit):
    for seed in range(N_GPU):
        rng = np.random.default_rng(9000 + seed)
        nf = int(rng.integers(1, 4))
        texts = [rand_unsorted(rng, int(rng.choice([0, 1, 2, 40, 300, 3000])), bool(rng.random() < 0.5)) for _ in range(nf)]
        want = O.sort_bed(texts)
        got = kit.sort_bed(b"".join(strip_headers(t) for t in texts))
        assert got == want, "seed %d" % seed


@pytest.mark.gpu
def test_device_sort_rejects_what_the_reference_rejects(kit):
    from bedops_b200._lib import BedKitError
    good = b"chr2\t5\t9\tz\nchr1\t1\t2\n"
    for bad in BAD_LINES:
        if not bad.endswith(b"\n") or bad.startswith(b"track"):
            continue   # the tool's business (unterminated last line, header zone)
        text = good + bad + good
        with pytest.raises(BedKitError) as ei:
            kit.sort_bed(text)
        assert ei.value.code == 4 and ei.value.bad_offset == len(good), bad[:40]


@pytest.mark.gpu
def test_device_sort_many_chromosomes_and_ties(kit):
    rng = np.random.default_rng(5)
    rows = []
    for k in range(200_000):
        c = "scaf%d" % int(rng.integers(0, 5000))
        s = int(rng.integers(0, 50))
        rows.append("%s\t%d\t%d\t%s" % (c, s, s + 1 + int(rng.integers(0, 3)), "r%d" % int(rng.integers(0, 1000))))
    rows += ["chrT\t5\t6\tsame%03d" % int(rng.integers(0, 400)) for _ in range(3000)]   # one long run of equal coordinates
    rng.shuffle(rows)
    text = ("\n".join(rows) + "\n").encode()
    assert kit.sort_bed(text) == O.sort_bed([text])


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_tool_matches_reference_binary(tmp_path):
    from bedops_b200._lib import tool_path
    ours = os.path.dirname(tool_path("sort-bed"))
    cases = []
    for seed in range(30):
        rng = np.random.default_rng(11000 + seed)
        nf = int(rng.integers(1, 4))
        files = {"f%d.bed" % k: rand_unsorted(rng, int(rng.choice([0, 3, 40, 300])), bool(rng.random() < 0.6)) for k in range(nf)}
        cases.append((sorted(files), files))
    good = b"chr2\t5\t9\tz\nchr1\t1\t2\n"
    for bad in BAD_LINES:
        cases.append((["a.bed", "b.bed"], {"a.bed": good, "b.bed": b"#h\n\n" + bad + (good if bad.endswith(b"\n") else b"")}))
    cases.append((["--max-mem", "1G", "--tmpdir", ".", "f.bed"], {"f.bed": b"chr1\t5\t6\nchr1\t1\t2\n"}))
    cases.append((["--max-mem", "10", "f.bed"], {"f.bed": b"chr1\t5\t6\n"}))
    cases.append((["--max-mem", "1T", "f.bed"], {"f.bed": b"chr1\t5\t6\n"}))
    cases.append((["--check-sort", "f.bed"], {"f.bed": b"chr1\t5\t6\nchr1\t1\t2\n"}))
    cases.append((["--check-sort", "f.bed"], {"f.bed": b"chr1\t1\t2\nchr1\t5\t6\n"}))
    cases.append((["nonexistent.bed"], {}))
    cases.append((["--check-sort", "nonexistent.bed"], {}))
    cases.append((["--help"], {}))
    cases.append((["--version"], {}))
    cases.append(([], {}))
    cases.append((["--max-mem"], {}))
    for argv, files in cases:
        assert run_ref("sort-bed", argv, files, tmp_path, ours) == run_ref("sort-bed", argv, files, tmp_path), argv
    # stdin
    data = b"chr1\t5\t6\tb\nchr1\t5\t6\ta\nchr1\t1\t2\n"
    a = subprocess.run([os.path.join(ours, "sort-bed"), "-"], input=data, capture_output=True)
    b = subprocess.run([os.path.join(REFBIN, "sort-bed"), "-"], input=data, capture_output=True)
    assert (a.returncode, a.stdout, a.stderr) == (b.returncode, b.stdout, b.stderr)


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="reference binaries not built")
def test_million_shuffled_rows_against_reference_binary(tmp_path):
    """BASELINE configs[0] scale: the 1 M-row synthetic map file, shuffled, sorted back -- must equal the reference
    binary's output byte for byte (and therefore the original file: it was sorted and has distinct rows)."""
    from bedops_b200 import synth
    from bedops_b200._lib import tool_path
    text = synth.bed_text(1_000_000, 1, synth.MAP_SHAPE)
    lines = text.split(b"\n")[:-1]
    rng = np.random.default_rng(3)
    rng.shuffle(lines)
    (tmp_path / "s.bed").write_bytes(b"\n".join(lines) + b"\n")
    ours = subprocess.run([tool_path("sort-bed"), "s.bed"], cwd=tmp_path, capture_output=True)
    ref = subprocess.run([os.path.join(REFBIN, "sort-bed"), "s.bed"], cwd=tmp_path, capture_output=True)
    assert ours.returncode == 0 and ours.stderr == b""
    assert ours.stdout == ref.stdout


@pytest.mark.gpu
def test_radix_sort_against_torch(kit):
    """the sorter's engine by itself: random keys of every width against torch's stable sort, keys only and with payload;
    sizes that leave partial tiles and sub-tiles, skewed digits (all equal, two values), 30 M keys for many tiles per CTA"""
    import torch
    g = torch.Generator(device="cuda:0")
    g.manual_seed(5)
    for n, nbits in [(2, 1), (33, 8), (2048, 16), (2049, 9), (16384, 64), (16385, 40), (100_003, 50), (1_000_000, 53), (30_000_000, 50)]:
        hi = (1 << min(nbits, 62)) - 1
        keys = torch.randint(0, hi + 1, (n,), device="cuda:0", dtype=torch.int64, generator=g)
        width = 8 * ((nbits + 7) // 8)          # whole 8-bit passes: the sort orders by key bits [0, width)
        if width <= 52:
            junk = torch.randint(0, 1 << 10, (n,), device="cuda:0", dtype=torch.int64, generator=g) << width   # bits the sort must ignore
            keys = keys | junk
        for skew in (0, 1, 2):
            k = keys.clone()
            if skew == 1:
                k[:] = k[0]
            elif skew == 2:
                k = torch.where(torch.arange(n, device="cuda:0") % 3 == 0, k[0], k[n // 2])
            vals = torch.arange(n, device="cuda:0", dtype=torch.int32)
            order = torch.sort(k & ((1 << width) - 1) if width < 63 else k, stable=True).indices
            torch.cuda.synchronize()
            kk, vv = k.clone(), vals.clone()
            kit.radix_sort_pairs(kk.data_ptr(), vv.data_ptr(), n, nbits)
            assert torch.equal(vv.long(), order), (n, nbits, skew)
            assert torch.equal(kk, k[order]), (n, nbits, skew)
            k2 = k.clone()
            kit.radix_sort_pairs(k2.data_ptr(), None, n, nbits)
            assert torch.equal(k2, k[order]), (n, nbits, skew)
