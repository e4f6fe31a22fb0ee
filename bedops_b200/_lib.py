"""ctypes binding of include/bedkit.h.  No compute happens in Python and there is no fallback: importing
the binding without a built libbedkit.so, or creating a BedKit without a B200, raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
BK_MAX_OPS = 64

OPS = {"echo": 1, "count": 2, "indicator": 3, "bases": 4, "sum": 5, "mean": 6, "max": 7, "min": 8,
       "echo-map-id": 9, "echo-ref-size": 10, "echo-ref-name": 11, "echo-ref-row-id": 12, "echo-map": 13,
       "echo-map-score": 14, "echo-map-size": 15, "echo-overlap-size": 16, "echo-map-range": 17, "bases-uniq": 18,
       "bases-uniq-f": 19, "variance": 20, "stdev": 21, "cv": 22,
       "echo-map-id-uniq": 23, "median": 24, "kth": 25, "mad": 26, "wmean": 27, "tmean": 28, "max-element": 29,
       "min-element": 30}
OVERLAP = {"bp": 0, "range": 1, "fraction-ref": 2, "fraction-map": 3, "fraction-either": 4, "fraction-both": 5,
           "exact": 6}
SETOPS = {"merge": 1, "intersect": 2, "element-of": 3, "not-element-of": 4, "complement": 5, "difference": 6,
          "symmdiff": 7, "everything": 8, "partition": 9}
COL_LINE, COL_SCORE, COL_ID, LOAD_HEADERS = 1, 2, 4, 8


class BedKitError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("bedkit error %d: %s" % (code, msg))
        self.code = code


class _Text(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_uint64), ("on_device", C.c_int), ("rows", C.c_uint64)]


class _MapSpec(C.Structure):
    _fields_ = [("n_ops", C.c_int), ("ops", C.c_int * BK_MAX_OPS), ("overlap_kind", C.c_int),
                ("overlap_bp", C.c_uint64), ("overlap_frac", C.c_double), ("precision", C.c_int), ("sci", C.c_int),
                ("skip_unmapped", C.c_int), ("delim", C.c_char_p), ("multidelim", C.c_char_p),
                ("chrom", C.c_char_p), ("out_on_device", C.c_int), ("op_arg", C.c_double * BK_MAX_OPS),
                ("row_id_base", C.c_uint64), ("op_arg2", C.c_double * BK_MAX_OPS)]


class _CfSpec(C.Structure):
    _fields_ = [("dist", C.c_int), ("closest", C.c_int), ("no_overlaps", C.c_int), ("no_ref", C.c_int),
                ("no_query", C.c_int), ("center", C.c_int), ("delim", C.c_char_p), ("chrom", C.c_char_p),
                ("out_on_device", C.c_int)]


def lib_path() -> str:
    # BEDKIT_LIB: an alternative build of the same library (kernel tuning experiments: profiles/tools/build_variants.sh)
    return os.environ.get("BEDKIT_LIB") or os.path.join(_HERE, "lib", "libbedkit.so")


def tool_path(name: str) -> str:
    return os.path.join(_HERE, "bin", name)


_LIB = None


def load_library() -> C.CDLL:
    """dlopen libbedkit.so and declare every prototype of include/bedkit.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError("libbedkit.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C bedops_b200/csrc`.  There is no CPU fallback." % path)
    lib = C.CDLL(path)
    vp, u64, i = C.c_void_p, C.c_uint64, C.c_int
    proto = {
        "bk_init": (i, [C.POINTER(vp), i]),
        "bk_destroy": (None, [vp]),
        "bk_set_stream": (i, [vp, vp]),
        "bk_sync": (i, [vp]),
        "bk_strerror": (C.c_char_p, [i]),
        "bk_last_error": (C.c_char_p, [vp]),
        "bk_abi_version": (i, []),
        "bk_launch_count": (u64, [vp]),
        "bk_profile": (i, [vp, i]),
        "bk_profile_query": (i, [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(u64)]),
        "bk_copy": (i, [vp, vp, vp, C.c_size_t]),
        "bk_release_cached": (i, [vp]),
        "bk_load_bed": (i, [vp, C.c_char_p, C.c_size_t, i, C.c_uint, C.POINTER(vp)]),
        "bk_load_bed_device": (i, [vp, vp, C.c_size_t, i, C.c_uint, C.POINTER(vp)]),
        "bk_free_bed": (None, [vp, vp]),
        "bk_bed_rows": (u64, [vp]),
        "bk_bed_nchrom": (i, [vp]),
        "bk_bed_chrom_name": (C.c_char_p, [vp, i]),
        "bk_bed_chrom_rows": (u64, [vp, i]),
        "bk_bed_copy_columns": (i, [vp, vp, vp, vp, vp, vp]),
        "bk_mapspec_default": (None, [C.POINTER(_MapSpec)]),
        "bk_bedmap": (i, [vp, vp, vp, C.POINTER(_MapSpec), C.POINTER(_Text)]),
        "bk_bedmap_host": (i, [vp, vp, C.c_size_t, i, C.c_uint, vp, C.c_size_t, i, C.c_uint, C.POINTER(_MapSpec),
                               C.POINTER(_Text)]),
        "bk_setop": (i, [vp, i, C.POINTER(vp), i, C.c_double, i, C.c_char_p, i, C.POINTER(_Text)]),
        "bk_chop": (i, [vp, C.POINTER(vp), i, u64, u64, i, C.c_char_p, i, C.POINTER(_Text)]),
        "bk_cfspec_default": (None, [C.POINTER(_CfSpec)]),
        "bk_closest": (i, [vp, vp, vp, C.POINTER(_CfSpec), C.POINTER(_Text)]),
        "bk_format_bed_device": (i, [vp, C.c_char_p, vp, vp, vp, u64, C.c_int64, C.POINTER(_Text)]),
        "bk_free_text": (None, [vp, C.POINTER(_Text)]),
        "bk_check_text": (i, [vp, C.c_char_p, C.c_size_t, i, i, i]),
        "bk_check_text_device": (i, [vp, vp, C.c_size_t, i, i, i]),
        "bk_find_start": (u64, [vp, u64, u64, u64]),
        "bk_bed_reach_start": (i, [vp, vp, C.c_char_p, u64, C.POINTER(u64)]),
        "bk_bed_chrom_max_end": (i, [vp, vp, C.c_char_p, C.POINTER(u64)]),
        "bk_bed_concat": (i, [vp, vp, vp, C.POINTER(vp)]),
        "bk_shard_free": (None, [vp, vp]),
        "bk_shard_bytes_in": (u64, [vp]),
        "bk_bedmap_shard_finish": (i, [vp, vp, C.POINTER(u64), C.POINTER(_Text)]),
        "bk_is_starch": (i, [C.c_char_p, C.c_size_t]),
        "bk_unstarch": (i, [vp, C.c_char_p, C.c_size_t, C.c_char_p, i, C.POINTER(_Text)]),
        "bk_starch_inflate_host": (i, [C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        "bk_host_free": (None, [vp]),
        "bk_radix_sort_pairs": (i, [vp, vp, vp, u64, i]),
        "bk_bed_pad": (i, [vp, vp, C.c_longlong, C.c_longlong, C.POINTER(vp)]),
        "bk_sort_bed": (i, [vp, C.c_char_p, C.c_size_t, i, C.POINTER(_Text), C.POINTER(u64)]),
        "bk_sort_bed_device": (i, [vp, vp, C.c_size_t, i, C.POINTER(_Text), C.POINTER(u64)]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


EXPORTS = ["bk_init", "bk_destroy", "bk_set_stream", "bk_sync", "bk_strerror", "bk_last_error", "bk_abi_version",
           "bk_launch_count", "bk_profile", "bk_profile_query", "bk_copy", "bk_load_bed", "bk_load_bed_device", "bk_free_bed", "bk_bed_rows", "bk_bed_nchrom",
           "bk_bed_chrom_name", "bk_bed_chrom_rows", "bk_bed_copy_columns", "bk_mapspec_default", "bk_bedmap",
           "bk_setop", "bk_cfspec_default", "bk_closest", "bk_format_bed_device", "bk_free_text", "bk_chrom_index", "bk_plan_shards", "bk_check_text",
           "bk_check_text_device", "bk_release_cached", "bk_bedmap_host", "bk_chop", "bk_find_start", "bk_plan_cuts",
           "bk_cut_offset", "bk_bed_reach_start", "bk_bed_chrom_max_end", "bk_bed_concat", "bk_shard_plan_make",
           "bk_bedmap_shard_begin", "bk_bedmap_shard_finish", "bk_shard_free", "bk_shard_bytes_in", "bk_sort_bed",
           "bk_sort_bed_device", "bk_bed_pad", "bk_is_starch", "bk_unstarch", "bk_starch_inflate_host",
           "bk_host_free", "bk_radix_sort_pairs"]


class Bed:
    """Handle of a device-resident parsed BED file."""

    def __init__(self, kit: "BedKit", handle: int, keep=None):
        self.kit, self.h, self._keep = kit, handle, keep

    @property
    def rows(self) -> int:
        return self.kit.lib.bk_bed_rows(self.h)

    def chroms(self) -> List[tuple]:
        lib = self.kit.lib
        return [(lib.bk_bed_chrom_name(self.h, k).decode(), lib.bk_bed_chrom_rows(self.h, k))
                for k in range(lib.bk_bed_nchrom(self.h))]

    def columns(self, score: bool = False, line: bool = False):
        import numpy as np
        n = self.rows
        st = np.empty(n, dtype=np.uint32)
        en = np.empty(n, dtype=np.uint32)
        sc = np.empty(n, dtype=np.float64) if score else None
        lo = np.empty(n, dtype=np.uint64) if line else None
        self.kit._chk(self.kit.lib.bk_bed_copy_columns(
            self.kit.ctx, self.h, st.ctypes.data, en.ctypes.data,
            sc.ctypes.data if score else None, lo.ctypes.data if line else None))
        return st, en, sc, lo

    def free(self):
        if self.h and getattr(self.kit, "ctx", None):
            self.kit.lib.bk_free_bed(self.kit.ctx, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceText:
    """Result text left in HBM (out_on_device)."""

    def __init__(self, kit, t: _Text):
        self.kit, self.t = kit, t

    @property
    def ptr(self) -> int:
        return self.t.ptr or 0

    @property
    def nbytes(self) -> int:
        return self.t.len

    @property
    def rows(self) -> int:
        return self.t.rows

    def free(self):
        if self.t.ptr and getattr(self.kit, "ctx", None):
            self.kit.lib.bk_free_text(self.kit.ctx, C.byref(self.t))
        self.t.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BedKit:
    """One engine context on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        ctx = C.c_void_p()
        rc = self.lib.bk_init(C.byref(ctx), device)
        if rc != 0:
            raise BedKitError(rc, self.lib.bk_strerror(rc).decode())
        self.ctx = ctx

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.bk_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int):
        if rc != 0:
            detail = self.lib.bk_last_error(self.ctx).decode()
            raise BedKitError(rc, detail or self.lib.bk_strerror(rc).decode())

    def set_stream(self, stream_ptr: Optional[int]):
        self._chk(self.lib.bk_set_stream(self.ctx, stream_ptr))

    def sync(self):
        self._chk(self.lib.bk_sync(self.ctx))

    @property
    def launches(self) -> int:
        return self.lib.bk_launch_count(self.ctx)

    def profile(self, enable: bool = True):
        """Enable (and reset) per-kernel CUDA-event timing inside the library."""
        self._chk(self.lib.bk_profile(self.ctx, int(enable)))

    def profile_query(self, kernel: str):
        """(total milliseconds, launches) of the named kernel since the last profile() reset."""
        ms, n = C.c_double(), C.c_uint64()
        self._chk(self.lib.bk_profile_query(self.ctx, kernel.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def release_cached(self):
        """hand the context's idle device blocks back to the driver"""
        self._chk(self.lib.bk_release_cached(self.ctx))

    def copy(self, dst_ptr: int, src_ptr: int, nbytes: int):
        self._chk(self.lib.bk_copy(self.ctx, dst_ptr, src_ptr, nbytes))

    # ---- reader -----------------------------------------------------------------------------------
    def load(self, text: bytes, min_fields: int = 3, cols: int = 0) -> Bed:
        h = C.c_void_p()
        self._chk(self.lib.bk_load_bed(self.ctx, text, len(text), min_fields, cols, C.byref(h)))
        return Bed(self, h)

    def load_host_ptr(self, ptr: int, nbytes: int, min_fields: int = 3, cols: int = 0) -> Bed:
        h = C.c_void_p()
        self._chk(self.lib.bk_load_bed(self.ctx, C.cast(ptr, C.c_char_p), nbytes, min_fields, cols, C.byref(h)))
        return Bed(self, h)

    def load_device(self, dev_ptr: int, nbytes: int, min_fields: int = 3, cols: int = 0, keep=None) -> Bed:
        h = C.c_void_p()
        self._chk(self.lib.bk_load_bed_device(self.ctx, dev_ptr, nbytes, min_fields, cols, C.byref(h)))
        return Bed(self, h, keep)

    def check_text(self, text: bytes, n_fields: int = 3, has_rest: bool = True, nest_check: bool = False):
        """--ec validation; raises BedKitError(code 9) with the reference's message text on the first bad line."""
        self._chk(self.lib.bk_check_text(self.ctx, text, len(text), n_fields, int(has_rest), int(nest_check)))

    # ---- tools ------------------------------------------------------------------------------------
    def free_text(self, t: _Text):
        self.lib.bk_free_text(self.ctx, C.byref(t))

    def _take(self, t: _Text, on_device: bool, raw: bool = False):
        if raw:
            return t          # caller frees with free_text(); host text stays in the library's pinned buffer
        if on_device:
            return DeviceText(self, t)
        # not C.string_at: its size argument is a C int, results can exceed 2 GiB
        data = bytes((C.c_char * t.len).from_address(t.ptr)) if t.len else b""
        self.lib.bk_free_text(self.ctx, C.byref(t))
        return data

    def bedmap_host(self, ref_ptr: int, ref_len: int, ref_fields: int, ref_cols: int, map_ptr: int, map_len: int,
                    map_fields: int, map_cols: int, ops: Sequence[str], _raw: bool = False, **kw):
        """bk_bedmap_host: the whole call over host buffers (addresses), transfers overlapped with the kernels."""
        spec = self._mapspec(ops, **kw)
        t = _Text()
        self._chk(self.lib.bk_bedmap_host(self.ctx, ref_ptr, ref_len, ref_fields, ref_cols, map_ptr, map_len, map_fields,
                                          map_cols, C.byref(spec), C.byref(t)))
        return self._take(t, False, _raw)

    def bedmap(self, ref: Bed, map_: Optional[Bed], ops: Sequence[str], overlap=("bp", 1), prec: int = 6,
               sci: bool = False, delim: bytes = b"|", multidelim: bytes = b";", skip_unmapped: bool = False,
               chrom: Optional[bytes] = None, on_device: bool = False, _raw: bool = False):
        spec = self._mapspec(ops, overlap, prec, sci, delim, multidelim, skip_unmapped, chrom, on_device)
        t = _Text()
        self._chk(self.lib.bk_bedmap(self.ctx, ref.h, map_.h if map_ is not None else None, C.byref(spec), C.byref(t)))
        return self._take(t, on_device, _raw)

    def _mapspec(self, ops: Sequence[str], overlap=("bp", 1), prec: int = 6, sci: bool = False, delim: bytes = b"|",
                 multidelim: bytes = b";", skip_unmapped: bool = False, chrom: Optional[bytes] = None,
                 on_device: bool = False):
        spec = _MapSpec()
        self.lib.bk_mapspec_default(C.byref(spec))
        spec.n_ops = len(ops)
        for k, o in enumerate(ops):
            name, _, arg = o.partition(":")          # "kth:0.25", "tmean:0.1:0.2"
            spec.ops[k] = OPS[name]
            a1, _, a2 = arg.partition(":")
            spec.op_arg[k] = float(a1) if a1 else 0.0
            spec.op_arg2[k] = float(a2) if a2 else 0.0
        kind, val = overlap
        spec.overlap_kind = OVERLAP[kind]
        if kind in ("bp", "range"):
            spec.overlap_bp = int(val)
        elif kind != "exact":
            spec.overlap_frac = float(val)
        spec.precision, spec.sci, spec.skip_unmapped = prec, int(sci), int(skip_unmapped)
        spec.delim, spec.multidelim, spec.chrom = delim, multidelim, chrom
        spec.out_on_device = int(on_device)
        self._keep_spec_strings = (delim, multidelim, chrom)
        return spec

    def setop(self, op: str, files: Sequence[Bed], thr: float = 1.0, use_pct: bool = True,
              chrom: Optional[bytes] = None, on_device: bool = False):
        arr = (C.c_void_p * len(files))(*[f.h for f in files])
        t = _Text()
        self._chk(self.lib.bk_setop(self.ctx, SETOPS[op], arr, len(files), thr, int(use_pct), chrom, int(on_device),
                                    C.byref(t)))
        return self._take(t, on_device)

    def chop(self, files: Sequence[Bed], chunk: int = 1, stagger: int = 0, exclude_short: bool = False,
             chrom: Optional[bytes] = None, on_device: bool = False):
        arr = (C.c_void_p * len(files))(*[f.h for f in files])
        t = _Text()
        self._chk(self.lib.bk_chop(self.ctx, arr, len(files), chunk, stagger, int(exclude_short), chrom, int(on_device),
                                   C.byref(t)))
        return self._take(t, on_device)

    def radix_sort_pairs(self, d_keys: int, d_vals: Optional[int], n: int, nbits: int):
        """stable radix sort of device arrays (u64 keys, optional u32 values) by key bits [0, nbits)"""
        self._chk(self.lib.bk_radix_sort_pairs(self.ctx, d_keys, d_vals, n, nbits))

    def unstarch(self, archive: bytes, chrom: Optional[bytes] = None, on_device: bool = False):
        """a Starch v2 archive -> the BED text `unstarch` prints"""
        t = _Text()
        self._chk(self.lib.bk_unstarch(self.ctx, archive, len(archive), chrom, int(on_device), C.byref(t)))
        return self._take(t, on_device)

    def pad(self, bed: "Bed", lpad: int, rpad: int) -> "Bed":
        """bedops --range L:R view of a loaded file (keeps `bed` alive: the view borrows its text)."""
        h = C.c_void_p()
        self._chk(self.lib.bk_bed_pad(self.ctx, bed.h, lpad, rpad, C.byref(h)))
        return Bed(self, h, keep=bed)

    def sort_bed(self, text: bytes, on_device: bool = False):
        """sort-bed over one text (files concatenated, leading headers removed, final NL present).  A row sort-bed
        rejects raises BedKitError with .bad_offset = byte offset of the offending line."""
        t = _Text()
        bad = C.c_uint64(0xFFFFFFFFFFFFFFFF)
        rc = self.lib.bk_sort_bed(self.ctx, text, len(text), int(on_device), C.byref(t), C.byref(bad))
        if rc != 0:
            err = BedKitError(rc, self.lib.bk_last_error(self.ctx).decode() or self.lib.bk_strerror(rc).decode())
            err.bad_offset = bad.value
            raise err
        return self._take(t, on_device)

    def sort_bed_device(self, dev_ptr: int, nbytes: int, on_device: bool = True):
        t = _Text()
        self._chk(self.lib.bk_sort_bed_device(self.ctx, dev_ptr, nbytes, int(on_device), C.byref(t), None))
        return self._take(t, on_device)

    def closest(self, ref: Bed, query: Bed, dist=False, closest=False, no_overlaps=False, no_ref=False,
                delim: bytes = b"|", chrom: Optional[bytes] = None, on_device: bool = False):
        spec = _CfSpec()
        self.lib.bk_cfspec_default(C.byref(spec))
        spec.dist, spec.closest, spec.no_overlaps, spec.no_ref = int(dist), int(closest), int(no_overlaps), int(no_ref)
        spec.delim, spec.chrom, spec.out_on_device = delim, chrom, int(on_device)
        t = _Text()
        self._chk(self.lib.bk_closest(self.ctx, ref.h, query.h, C.byref(spec), C.byref(t)))
        return self._take(t, on_device)

    def format_bed_device(self, chrom: bytes, d_start: int, d_end: int, d_score: int, n: int, id_base: int) -> DeviceText:
        t = _Text()
        self._chk(self.lib.bk_format_bed_device(self.ctx, chrom, d_start, d_end, d_score, n, id_base, C.byref(t)))
        return DeviceText(self, t)
