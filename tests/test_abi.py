"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/bedkit.h declares,
has no torch types in its ABI, and the product fails loudly (no CPU fallback) when there is no B200."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT, REFBIN, have_ref


def header_symbols():
    src = open(os.path.join(ROOT, "include", "bedkit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bk_[a-z_0-9]+)\s*\(", src)))


def lib_file():
    import bedops_b200
    return bedops_b200.lib_path()


def gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except ImportError:
        return False


def test_library_exports_every_declared_symbol():
    import bedops_b200
    lib = bedops_b200.load_library()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libbedkit.so does not export %s" % s
    from bedops_b200._lib import EXPORTS
    assert sorted(EXPORTS) == syms
    assert lib.bk_abi_version() == 3


def test_abi_is_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", lib_file()], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    bk = [n for n in names if n.startswith("bk_")]
    assert len(bk) >= 20
    assert not any("torch" in n or "at::" in n for n in names)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", lib_file()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import bedops_b200
    if gpu_present():
        pytest.skip("a GPU is present")
    with pytest.raises(bedops_b200.BedKitError) as e:
        bedops_b200.BedKit()
    assert e.value.code == 1
    for tool in ("bedmap", "bedops", "closest-features"):
        assert os.access(bedops_b200.tool_path(tool), os.X_OK)
    p = subprocess.run([bedops_b200.tool_path("bedops"), "-m", os.path.join(ROOT, "tests", "golden", "docs.json")],
                       capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout == "" and "no CPU fallback" in p.stderr


def test_product_does_not_import_oracle():
    for dirpath, _, names in os.walk(os.path.join(ROOT, "bedops_b200")):
        for n in names:
            if n.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, n), errors="ignore").read()
                assert "bed_oracle" not in src and "oracle/" not in src, n


def test_cli_banners_and_errors_match_reference_shape():
    import bedops_b200
    p = subprocess.run([bedops_b200.tool_path("bedops"), "--version"], capture_output=True, text=True)
    assert p.returncode == 0
    assert p.stdout == ("bedops\n  citation: http://bioinformatics.oxfordjournals.org/content/28/14/1919.abstract\n"
                        "  version:  2.4.26\n  authors:  Shane Neph & Scott Kuehn\n")
    p = subprocess.run([bedops_b200.tool_path("bedmap"), "--version"], capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout.startswith("bedmap\n  citation:")   # Bedmap.cpp:166-170 quirk
    p = subprocess.run([bedops_b200.tool_path("bedops"), "-m", "/nonexistent.bed"], capture_output=True, text=True)
    assert p.returncode == 1
    assert p.stderr == "May use bedops --help for more help.\n\nError: Bad Input\nCannot find /nonexistent.bed\n"
    p = subprocess.run([bedops_b200.tool_path("bedmap"), "--count", "/nonexistent.bed"], capture_output=True, text=True)
    assert p.returncode == 1
    assert p.stderr == "May use bedmap --help for more help.\n\nError: Unable to find file: /nonexistent.bed\n"
    p = subprocess.run([bedops_b200.tool_path("bedmap"), "--bogus", "a", "b"], capture_output=True, text=True)
    assert p.stderr == "May use bedmap --help for more help.\n\nError: Unknown option: --bogus\n"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/bin not built")
def test_help_and_usage_texts_are_the_reference_s_byte_for_byte():
    """--help (stdout, exit 0), no arguments (stderr, exit 1), bedops --help-<operation>: part of the process-boundary
    contract.  Needs no GPU: the tools print them before touching the library."""
    import bedops_b200
    for tool in ("bedmap", "bedops", "closest-features"):
        for argv in (["--help"], []):
            exp = subprocess.run([os.path.join(REFBIN, tool)] + argv, capture_output=True)
            got = subprocess.run([bedops_b200.tool_path(tool)] + argv, capture_output=True)
            assert (got.returncode, got.stdout, got.stderr) == (exp.returncode, exp.stdout, exp.stderr), (tool, argv)
    for op in ("merge", "intersect", "element-of", "not-element-of", "complement", "difference", "symmdiff", "chop",
               "everything", "partition"):
        exp = subprocess.run([os.path.join(REFBIN, "bedops"), "--help-" + op], capture_output=True)
        got = subprocess.run([bedops_b200.tool_path("bedops"), "--help-" + op], capture_output=True)
        assert (got.returncode, got.stdout) == (exp.returncode, exp.stdout), op


def test_more_operations_than_the_spec_table_holds_is_a_clean_error(tmp_path):
    """200 x --count used to smash the stack of the tool (bk_mapspec.ops is a fixed table)."""
    import bedops_b200
    (tmp_path / "a.bed").write_bytes(b"chr1\t1\t2\n")
    p = subprocess.run([bedops_b200.tool_path("bedmap")] + ["--count"] * 200 + ["a.bed", "a.bed"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and "operations given" in p.stderr and "stack smashing" not in p.stderr


def test_entry_points_answer_null_handles_with_an_error_code_not_a_crash():
    """No GPU needed: every entry point that takes a ctx / bed / shard / text looks at its pointers before it touches the
    device (the two *_default functions fill a caller-provided struct and are the exception, as in any C API)."""
    import bedops_b200
    bedops_b200.load_library()
    calls = {
      "bk_destroy": "lib.bk_destroy(None)",
      "bk_set_stream": "lib.bk_set_stream(None,None)",
      "bk_sync": "lib.bk_sync(None)",
      "bk_last_error": "lib.bk_last_error(None)",
      "bk_launch_count": "lib.bk_launch_count(None)",
      "bk_profile": "lib.bk_profile(None,1)",
      "bk_profile_query": "lib.bk_profile_query(None,b'k',None,None)",
      "bk_copy": "lib.bk_copy(None,None,None,C.c_size_t(0))",
      "bk_release_cached": "lib.bk_release_cached(None)",
      "bk_load_bed": "lib.bk_load_bed(None,b'x',C.c_size_t(1),3,0,None)",
      "bk_load_bed_device": "lib.bk_load_bed_device(None,None,C.c_size_t(0),3,0,None)",
      "bk_free_bed": "lib.bk_free_bed(None,None)",
      "bk_bed_rows": "lib.bk_bed_rows(None)",
      "bk_bed_nchrom": "lib.bk_bed_nchrom(None)",
      "bk_bed_chrom_name": "lib.bk_bed_chrom_name(None,0)",
      "bk_bed_chrom_rows": "lib.bk_bed_chrom_rows(None,0)",
      "bk_bed_copy_columns": "lib.bk_bed_copy_columns(None,None,None,None,None,None)",
      "bk_check_text": "lib.bk_check_text(None,b'x',C.c_size_t(1),3,1,0)",
      "bk_check_text_device": "lib.bk_check_text_device(None,None,C.c_size_t(0),3,1,0)",
      "bk_bedmap": "lib.bk_bedmap(None,None,None,None,None)",
      "bk_bedmap_host": "lib.bk_bedmap_host(None,None,C.c_size_t(0),3,0,None,C.c_size_t(0),3,0,None,None)",
      "bk_setop": "lib.bk_setop(None,0,None,0,C.c_double(1.0),0,None,0,None)",
      "bk_chop": "lib.bk_chop(None,None,0,C.c_uint64(1),C.c_uint64(0),0,None,0,None)",
      "bk_bed_pad": "lib.bk_bed_pad(None,None,C.c_longlong(0),C.c_longlong(0),None)",
      "bk_closest": "lib.bk_closest(None,None,None,None,None)",
      "bk_is_starch": "lib.bk_is_starch(None,C.c_size_t(0))",
      "bk_unstarch": "lib.bk_unstarch(None,None,C.c_size_t(0),None,0,None)",
      "bk_starch_inflate_host": "lib.bk_starch_inflate_host(None,C.c_size_t(0),None,None,None)",
      "bk_host_free": "lib.bk_host_free(None)",
      "bk_sort_bed": "lib.bk_sort_bed(None,None,C.c_size_t(0),0,None,None)",
      "bk_sort_bed_device": "lib.bk_sort_bed_device(None,None,C.c_size_t(0),0,None,None)",
      "bk_radix_sort_pairs": "lib.bk_radix_sort_pairs(None,None,None,C.c_uint64(0),64)",
      "bk_format_bed_device": "lib.bk_format_bed_device(None,None,None,None,None,None,C.c_uint64(0),None)",
      "bk_free_text": "lib.bk_free_text(None,None)",
      "bk_chrom_index": "lib.bk_chrom_index(None,C.c_size_t(0),None,0,None)",
      "bk_plan_shards": "lib.bk_plan_shards(None,0,0,None)",
      "bk_find_start": "lib.bk_find_start(None,C.c_uint64(0),C.c_uint64(10),C.c_uint64(1))",
      "bk_plan_cuts": "lib.bk_plan_cuts(None,C.c_size_t(0),None,0,0,None)",
      "bk_cut_offset": "lib.bk_cut_offset(None,C.c_size_t(0),None,0,None)",
      "bk_bed_reach_start": "lib.bk_bed_reach_start(None,None,None,C.c_uint64(0),None)",
      "bk_bed_chrom_max_end": "lib.bk_bed_chrom_max_end(None,None,None,None)",
      "bk_bed_concat": "lib.bk_bed_concat(None,None,None,None)",
      "bk_shard_plan_make": "lib.bk_shard_plan_make(None,C.c_size_t(0),None,C.c_size_t(0),0,None)",
      "bk_bedmap_shard_begin": "lib.bk_bedmap_shard_begin(None,None,0,None,C.c_size_t(0),3,0,None,C.c_size_t(0),3,0,None,None,None)",
      "bk_bedmap_shard_finish": "lib.bk_bedmap_shard_finish(None,None,None,None)",
      "bk_shard_free": "lib.bk_shard_free(None,None)",
      "bk_shard_bytes_in": "lib.bk_shard_bytes_in(None)",
    }
    prog = ("import ctypes as C\nlib = C.CDLL(%r)\nlib.bk_last_error.restype = C.c_char_p\nlib.bk_bed_chrom_name.restype = C.c_char_p\n"
            "lib.bk_find_start.restype = C.c_uint64\nlib.bk_cut_offset.restype = C.c_uint64\n" % bedops_b200.lib_path())
    for name, expr in calls.items():
        prog += "r = %s\nprint(%r, r)\n" % (expr, name)
    p = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, (p.stdout[-300:], p.stderr[-300:])
    seen = dict(line.split(" ", 1) for line in p.stdout.strip().split("\n"))
    assert len(seen) == len(calls)
    for name in ("bk_load_bed", "bk_bedmap", "bk_bedmap_host", "bk_setop", "bk_closest", "bk_sort_bed", "bk_unstarch", "bk_bed_pad",
                 "bk_bedmap_shard_begin", "bk_bedmap_shard_finish", "bk_check_text", "bk_chop", "bk_bed_concat"):
        assert seen[name] == "3", (name, seen[name])   # BK_ERR_ARG
