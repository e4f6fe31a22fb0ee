"""Starch v2 archives as input.  CPU: the Python restatement of the reader against the reference's own
`unstarch` on archives made by the reference's `starch` (bzip2 and gzip), the committed golden archives, and the library's
host stage (container walk + inflate) against the restatement.  GPU: bk_unstarch (device un-transform) against both, and
the drop-in tools reading archives against the reference tools reading the same archives."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, REFBIN, have_ref
import bed_oracle as O
from test_random_differential import rand_bed


def have_starch():
    return os.access(os.path.join(REFBIN, "starch"), os.X_OK) and os.access(os.path.join(REFBIN, "unstarch"), os.X_OK)


def make_archive(tmp_path, text, flag):
    (tmp_path / "in.bed").write_bytes(text)
    return subprocess.run([os.path.join(REFBIN, "starch"), flag, "in.bed"], cwd=tmp_path, capture_output=True, check=True).stdout


def rand_sorted_bed(rng):
    chroms = [["chr1"], ["chr1", "chr2"], ["chr1", "chr10", "chr2", "chrX", "scaffold_12"]][int(rng.integers(0, 3))]
    t = rand_bed(rng, int(rng.choice([1, 5, 60, 400])), int(rng.choice([60, 5000, 200_000_000])), chroms, fields=int(rng.choice([3, 5])))
    rows = O.parse_bed(t, 3)
    rows.sort(key=lambda r: (r.chrom, r.start, r.end, r.rest3))
    return b"".join(O.echo_b3rest(r) + b"\n" for r in rows)


def host_inflate(archive, chrom=None):
    import bedops_b200
    lib = bedops_b200.load_library()
    t, n = C.c_void_p(), C.c_size_t()
    rc = lib.bk_starch_inflate_host(archive, len(archive), chrom, C.byref(t), C.byref(n))
    if rc != 0:
        return rc, None
    data = C.string_at(t.value, n.value)
    lib.bk_host_free(t)
    return 0, data


def test_golden_archives():
    want = open(os.path.join(GOLDEN, "starch_expected.bed"), "rb").read()
    for name in ("starch_bz2.starch", "starch_gz.starch"):
        blob = open(os.path.join(GOLDEN, name), "rb").read()
        assert O.unstarch(blob) == want
        rc, got = host_inflate(blob)
        assert rc == 0 and got == b"".join(b">" + c + b"\n" + x for c, x in O.starch_streams(blob))
        rc, got = host_inflate(blob, b"chr10")
        assert rc == 0 and got == b"".join(b">" + c + b"\n" + x for c, x in O.starch_streams(blob) if c == b"chr10")
    assert host_inflate(b"BZh91AY&SY" + b"\0" * 200)[0] == 7      # a v1 archive / bare bzip2 stream: BK_ERR_STARCH
    assert host_inflate(blob[:4] + b"\0" * 300)[0] == 7             # magic without a footer


def test_damaged_archives_are_refused_not_followed():
    """The host stage walks offsets and sizes it reads from the archive: truncated files, a footer that points outside,
    stream sizes that overrun (or would wrap a 64-bit sum) and flipped bytes all end in BK_ERR_STARCH or in a clean
    inflate, never in a read outside the buffer."""
    import random
    import re
    rng = random.Random(11)
    for name in ("starch_bz2.starch", "starch_gz.starch"):
        blob = open(os.path.join(GOLDEN, name), "rb").read()
        for cut in (4, 5, 130, 131, len(blob) // 2, len(blob) - 127, len(blob) - 1):
            assert host_inflate(blob[:cut])[0] == 7, cut
        for off in (0, 3, len(blob) - 126, len(blob), 2 ** 63, 10 ** 19):
            b = bytearray(blob)
            b[len(b) - 127:len(b) - 107] = (b"%020d" % off)[:20]
            assert host_inflate(bytes(b))[0] == 7, off
        m = re.search(rb'"size"\s*:\s*"?(\d+)', blob)
        for size in (b"18446744073709551615", b"18446744073709551000", b"99999999999999999999999", b"%d" % len(blob)):
            b = blob[:m.start(1)] + size + blob[m.end(1):]   # the metadata follows the streams: its offset in the footer still holds
            assert host_inflate(b)[0] == 7, size
        for _ in range(150):
            b = bytearray(blob)
            for _ in range(rng.randrange(1, 6)):
                b[rng.randrange(4, len(b))] = rng.randrange(256)
            assert host_inflate(bytes(b))[0] in (0, 7)


@pytest.mark.skipif(not have_starch(), reason="reference starch/unstarch not built")
def test_oracle_matches_reference_unstarch(tmp_path):
    for seed in range(60):
        rng = np.random.default_rng(80000 + seed)
        text = rand_sorted_bed(rng)
        if not text:
            continue
        blob = make_archive(tmp_path, text, "--gzip" if seed % 2 else "--bzip2")
        (tmp_path / "a.starch").write_bytes(blob)
        ref = subprocess.run([os.path.join(REFBIN, "unstarch"), "a.starch"], cwd=tmp_path, capture_output=True, check=True).stdout
        assert ref == text == O.unstarch(blob), seed
        rc, got = host_inflate(blob)
        assert rc == 0 and got == b"".join(b">" + c + b"\n" + x for c, x in O.starch_streams(blob)), seed


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def kit():
    import bedops_b200
    k = bedops_b200.BedKit(0)
    yield k
    k.close()


@pytest.mark.gpu
def test_device_unstarch_golden(kit):
    want = open(os.path.join(GOLDEN, "starch_expected.bed"), "rb").read()
    for name in ("starch_bz2.starch", "starch_gz.starch"):
        blob = open(os.path.join(GOLDEN, name), "rb").read()
        assert kit.unstarch(blob) == want
        assert kit.unstarch(blob, b"chr10") == b"".join(l + b"\n" for l in want.split(b"\n") if l.startswith(b"chr10\t"))
        assert kit.unstarch(blob, b"nope") == b""


@pytest.mark.gpu
@pytest.mark.skipif(not have_starch(), reason="reference starch/unstarch not built")
def test_device_unstarch_random(kit, tmp_path):
    for seed in range(80):
        rng = np.random.default_rng(81000 + seed)
        text = rand_sorted_bed(rng)
        if not text:
            continue
        blob = make_archive(tmp_path, text, "--gzip" if seed % 2 else "--bzip2")
        assert kit.unstarch(blob) == text, seed


@pytest.mark.gpu
@pytest.mark.skipif(not (have_ref() and have_starch()), reason="reference binaries not built")
def test_tools_read_archives_like_the_reference(tmp_path):
    from bedops_b200 import synth
    from bedops_b200._lib import tool_path
    ref_t = synth.bed_text(2000, 2, synth.REF_SHAPE)
    map_t = synth.bed_text(20000, 1, synth.MAP_SHAPE)
    (tmp_path / "r.bed").write_bytes(ref_t)
    (tmp_path / "m.bed").write_bytes(map_t)
    (tmp_path / "r.starch").write_bytes(make_archive(tmp_path, ref_t, "--bzip2"))
    (tmp_path / "m.starch").write_bytes(make_archive(tmp_path, map_t, "--gzip"))
    cases = [("bedmap", ["--echo", "--count", "--mean", "--bases", "r.starch", "m.starch"]),
             ("bedmap", ["--echo", "--echo-map-id", "r.bed", "m.starch"]),
             ("bedops", ["-m", "r.starch", "m.starch"]),
             ("bedops", ["-e", "1", "r.starch", "m.bed"]),
             ("bedops", ["-u", "r.starch", "m.starch"]),
             ("closest-features", ["--dist", "r.starch", "m.starch"])]
    for tool, argv in cases:
        exp = subprocess.run([os.path.join(REFBIN, tool)] + argv, cwd=tmp_path, capture_output=True)
        got = subprocess.run([tool_path(tool)] + argv, cwd=tmp_path, capture_output=True)
        assert (got.returncode, got.stdout) == (exp.returncode, exp.stdout), (tool, argv, got.stderr[:300])
        assert exp.returncode == 0 and len(exp.stdout) > 100


@pytest.mark.gpu
@pytest.mark.skipif(not have_starch(), reason="reference starch/unstarch not built")
def test_device_unstarch_million_rows(kit, tmp_path):
    from bedops_b200 import synth
    text = synth.bed_text(1_000_000, 1, synth.MAP_SHAPE)
    blob = make_archive(tmp_path, text, "--bzip2")
    assert kit.unstarch(blob) == text
