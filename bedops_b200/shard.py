"""Genomic-range sharding of sorted BED inputs across GPUs -- thin host-side wrapper over the C ABI's
planner (bk_chrom_index / bk_plan_shards, csrc/hostplan.cu).  No compute happens here.

The path shards naturally: every chromosome is independent, so a shard is a contiguous group of chromosomes (in
strcmp order, which is file order), each file contributes one contiguous byte slice per shard, every shard is an
ordinary single-GPU call, and the outputs concatenated in shard order are byte-identical to the unsharded run --
the same contract as the reference's one-process-per-chromosome scale-out (bedmap/src/Input.hpp:117-122)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

from ._lib import load_library


class _Span(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("begin", C.c_uint64), ("end", C.c_uint64)]


def chrom_index(text: bytes) -> List[Tuple[bytes, int, int]]:
    """[(chromosome, byte_begin, byte_end)] of a sorted BED text."""
    lib = load_library()
    lib.bk_chrom_index.restype = C.c_int
    lib.bk_chrom_index.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Span), C.c_int, C.POINTER(C.c_int)]
    n = C.c_int(0)
    cap = 64
    while True:
        arr = (_Span * cap)()
        rc = lib.bk_chrom_index(text, len(text), arr, cap, C.byref(n))
        if rc == 0:
            return [(arr[i].name, arr[i].begin, arr[i].end) for i in range(n.value)]
        if rc != 2:
            raise RuntimeError("bk_chrom_index failed: %d" % rc)
        cap = n.value + 8


def plan_shards(files: Sequence[bytes], n_shards: int):
    """Balanced contiguous partition of the chromosomes of `files` (load = bytes over all files).
    Returns [{"chroms": [...], "slices": [(begin, end) per file]}] of length n_shards (shards may be empty)."""
    lib = load_library()
    lib.bk_plan_shards.restype = C.c_int
    lib.bk_plan_shards.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_int, C.POINTER(C.c_int)]
    idx = [chrom_index(t) for t in files]
    names = sorted({c for ix in idx for c, _, _ in ix})
    pos = {c: k for k, c in enumerate(names)}
    load = [0] * len(names)
    for ix in idx:
        for c, b, e in ix:
            load[pos[c]] += e - b
    first = (C.c_int * (n_shards + 1))()
    arr = (C.c_uint64 * max(1, len(names)))(*load)
    rc = lib.bk_plan_shards(arr, len(names), n_shards, first)
    if rc != 0:
        raise RuntimeError("bk_plan_shards failed: %d" % rc)
    shards = []
    for s in range(n_shards):
        group = names[first[s]:first[s + 1]]
        gs = set(group)
        slices = []
        for ix in idx:
            spans = [(b, e) for c, b, e in ix if c in gs]
            slices.append((spans[0][0], spans[-1][1]) if spans else (0, 0))
        shards.append({"chroms": group, "slices": slices, "load": sum(load[first[s]:first[s + 1]])})
    return shards


# ---- cuts inside chromosomes: ONE dataset over N GPUs by balanced genomic ranges with halos (csrc/hostplan.cu, shard.cu,
# pipeline.cu; protocol described in include/bedkit.h) ---------------------------------------------------------------------
BK_MAX_SHARDS = 64
INF = (1 << 64) - 1


class Cut(C.Structure):
    _fields_ = [("chrom", C.c_char * 128), ("coord", C.c_uint64), ("at_end", C.c_int)]


class ShardPlan(C.Structure):
    _fields_ = [("n_shards", C.c_int), ("cuts", Cut * BK_MAX_SHARDS), ("ref_off", C.c_uint64 * (BK_MAX_SHARDS + 1)),
                ("map_off", C.c_uint64 * (BK_MAX_SHARDS + 1)), ("map_chrom_begin", C.c_uint64 * BK_MAX_SHARDS),
                ("map_chrom_end", C.c_uint64 * BK_MAX_SHARDS)]


def _addr(buf):
    """(address, nbytes) of bytes / bytearray / numpy array / mmap / (address, nbytes)."""
    if isinstance(buf, tuple):
        return int(buf[0]), int(buf[1])
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value or 0, len(buf)
    import numpy as np
    a = np.frombuffer(buf, dtype=np.uint8)
    return a.ctypes.data, a.size


def find_start(text, begin: int, end: int, coord: int) -> int:
    """byte offset of the first record of [begin,end) (one chromosome) whose start >= coord (bk_find_start)"""
    lib = load_library()
    p, _ = _addr(text)
    return lib.bk_find_start(p, begin, end, coord)


def make_plan(ref_text, map_text, n_shards: int) -> ShardPlan:
    """bk_shard_plan_make: byte-balanced genomic ranges of the larger file, located in the other one (host only)."""
    lib = load_library()
    lib.bk_shard_plan_make.restype = C.c_int
    lib.bk_shard_plan_make.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(ShardPlan)]
    plan = ShardPlan()
    (rp, rn), (mp, mn) = _addr(ref_text), _addr(map_text)
    rc = lib.bk_shard_plan_make(rp, rn, mp, mn, n_shards, C.byref(plan))
    if rc != 0:
        raise RuntimeError("bk_shard_plan_make failed: %d" % rc)
    return plan


class RangeShard:
    """One rank's half-finished range-sharded bedmap call: begin() -> exchange reach vectors -> finish()."""

    def __init__(self, kit, plan: ShardPlan, rank: int, ref_text, map_text, ops, ref_fields=3, ref_cols=0, map_fields=5,
                 map_cols=0, ref_src=None, map_src=None, on_device=False, **spec_kw):
        from ._lib import _Text  # noqa: F401
        self.kit, self.plan, self.rank = kit, plan, rank
        lib = kit.lib
        lib.bk_bedmap_shard_begin.restype = C.c_int
        lib.bk_bedmap_shard_begin.argtypes = [C.c_void_p, C.POINTER(ShardPlan), C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_uint,
                                              C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        self._keep = (ref_text, map_text, ref_src, map_src)
        (rp, rn), (mp, mn) = _addr(ref_text), _addr(map_text)
        spec = kit._mapspec(ops, on_device=on_device, **spec_kw)   # where the result text is left: HBM or pinned host
        self._spec = spec
        self.on_device = on_device
        self.h = C.c_void_p()
        self.reach = (C.c_uint64 * plan.n_shards)()
        kit._chk(lib.bk_bedmap_shard_begin(kit.ctx, C.byref(plan), rank, rp, rn, ref_fields, ref_cols, mp, mn, map_fields, map_cols,
                                           _addr(ref_src)[0] if ref_src is not None else None,
                                           _addr(map_src)[0] if map_src is not None else None, C.byref(spec), C.byref(self.h),
                                           self.reach))

    def reach_list(self):
        return [int(x) for x in self.reach]

    def finish(self, all_reach, _raw=False):
        """all_reach: n_shards x n_shards (row i = rank i's reach list).  Returns this rank's output text (_raw: the
        library's own pinned bk_text, not copied into a Python object; release with kit.free_text)."""
        from ._lib import _Text
        n = self.plan.n_shards
        flat = (C.c_uint64 * (n * n))(*[int(v) for row in all_reach for v in row])
        t = _Text()
        h, self.h = self.h, None
        try:
            self.kit._chk(self.kit.lib.bk_bedmap_shard_finish(self.kit.ctx, h, flat, C.byref(t)))
            self.bytes_in = self.kit.lib.bk_shard_bytes_in(h)   # own slices + halos copied from the source
        finally:
            self.kit.lib.bk_shard_free(self.kit.ctx, h)
        return self.kit._take(t, self.on_device, _raw)

    def __del__(self):
        try:
            if self.h:
                self.kit.lib.bk_shard_free(self.kit.ctx, self.h)
        except Exception:
            pass
