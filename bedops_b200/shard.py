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
