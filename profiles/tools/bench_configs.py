#!/usr/bin/env python3
"""Device-resident timings of BASELINE.json configurations 3-5 (bench.py covers configuration 2).

One JSON line per operation: wall time by CUDA events around the C-ABI calls (text already in HBM, result left in HBM)
and the per-kernel split from the library's own event pairs (bk_profile).  Not a bench line -- supporting numbers for
DESIGN.md section 6."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KERNELS = ["k_fill_ids", "k_sort_validate", "k_chrom_insert", "k_sort_keys", "k_radix_hist", "k_radix_scatter", "k_tie_fix", "k_part_keys",
           "k_efflen", "k_count_rows", "k_scan_warps", "k_parse", "k_pmax_reduce", "k_pmax", "k_rank_merge", "k_segments",
           "k_intersect", "k_element_of", "k_argmark", "k_cf_sim", "k_map_group", "k_map_stats", "k_emit_len", "k_emit"]


def main():
    import torch
    import bedops_b200
    from bedops_b200._lib import COL_LINE, COL_SCORE
    from test_gpu_scale import SynthFile, MAP_SHAPE, REF_SHAPE
    scale = float(os.environ.get("BEDKIT_SCALE", "1.0"))
    kit = bedops_b200.BedKit(0)
    torch.cuda.set_device(0)

    def timed(name, units, fn, reps=3):
        fn()  # warm-up (also sizes the memory pool)
        kit.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        split = {}
        for k in KERNELS:
            t, n = kit.profile_query(k)
            if n:
                split[k] = round(t / reps, 4)
        kit.profile(False)
        print(json.dumps({"op": name, "input_rows": units, "ms": round(ms, 3), "rows_per_s": units / (ms * 1e-3),
                          "kernel_ms": split}), flush=True)

    only = os.environ.get("BEDKIT_CONFIGS", "3,4,5").split(",")
    if "2ids" in only:
      run_config2_ids(kit, torch, timed, scale, SynthFile, MAP_SHAPE, REF_SHAPE)
    if "sort" in only:
      run_sort(kit, torch, timed, scale, SynthFile, MAP_SHAPE)
    if "3" in only:
      run_config3(kit, torch, timed, scale, SynthFile, MAP_SHAPE, COL_LINE)
    if "4" in only:
      run_config4(kit, torch, timed, scale, SynthFile, MAP_SHAPE, REF_SHAPE, COL_LINE)
    if "5" in only:
      run_config5(kit, torch, timed, scale, SynthFile, MAP_SHAPE, REF_SHAPE, COL_LINE, COL_SCORE)
    kit.close()


def run_config2_ids(kit, torch, timed, scale, SynthFile, MAP_SHAPE, REF_SHAPE):
    # configuration 2 with the list operation north_star names: --echo-map-id (about 70 ids per reference row)
    from bedops_b200._lib import COL_ID, COL_LINE
    ref = SynthFile(kit, torch, int(10_000_000 * scale), 2, REF_SHAPE)
    mp = SynthFile(kit, torch, int(100_000_000 * scale), 1, MAP_SHAPE)

    def ids():
        rb, mb = ref.load(kit, 3, COL_LINE), mp.load(kit, 4, COL_ID | COL_LINE)
        out = kit.bedmap(rb, mb, ["echo", "echo-map-id"], on_device=True)
        ids.bytes = out.nbytes
        out.free()
        rb.free()
        mb.free()
    timed("bedmap --echo --echo-map-id, 10M x 100M", ref.rows + mp.rows, ids)
    print(json.dumps({"op": "bedmap --echo --echo-map-id, 10M x 100M", "out_bytes": ids.bytes}), flush=True)
    del ref, mp
    kit.release_cached()
    torch.cuda.empty_cache()


def run_sort(kit, torch, timed, scale, SynthFile, MAP_SHAPE):
    # sort-bed over a sorted 100 M-row file (the radix sort does the same passes whatever the input order) and
    # bedops --partition over it
    f = SynthFile(kit, torch, int(100_000_000 * scale), 1, MAP_SHAPE)

    def sort():
        out = kit.sort_bed_device(f.buf.data_ptr(), f.nbytes, on_device=True)
        out.free()
    timed("sort-bed, 100M rows", f.rows, sort)

    def part():
        b = f.load(kit, 3, 0)
        out = kit.setop("partition", [b], on_device=True)
        out.free()
        b.free()
    timed("bedops --partition, 100M rows", f.rows, part)
    del f
    kit.release_cached()
    torch.cuda.empty_cache()


def run_config3(kit, torch, timed, scale, SynthFile, MAP_SHAPE, COL_LINE):
    # configuration 3: bedops over 4 files of 250 M rows (load + operation, as the tools do)
    files = [SynthFile(kit, torch, int(250_000_000 * scale), s, MAP_SHAPE) for s in (1, 3, 4, 5)]
    rows = sum(f.rows for f in files)

    def setop(op, thr=1.0, pct=True):
        def run():
            beds = [f.load(kit, 3, COL_LINE if (k == 0 and "element" in op) else 0) for k, f in enumerate(files)]
            out = kit.setop(op, beds, thr, pct, on_device=True)
            out.free()
            for b in beds:
                b.free()
        return run
    timed("bedops --merge, 4 x 250M", rows, setop("merge"))
    timed("bedops --intersect, 4 x 250M", rows, setop("intersect"))
    timed("bedops --element-of 1, 4 x 250M", rows, setop("element-of", 1, False))
    timed("bedops --not-element-of 100%, 4 x 250M", rows, setop("not-element-of", 1.0, True))
    del files
    kit.release_cached()
    torch.cuda.empty_cache()



def run_config4(kit, torch, timed, scale, SynthFile, MAP_SHAPE, REF_SHAPE, COL_LINE):
    # configuration 4: closest-features 50 M x 200 M
    ref = SynthFile(kit, torch, int(50_000_000 * scale), 2, REF_SHAPE)
    qry = SynthFile(kit, torch, int(200_000_000 * scale), 1, MAP_SHAPE)

    def closest():
        rb, qb = ref.load(kit, 3, COL_LINE), qry.load(kit, 3, COL_LINE)
        out = kit.closest(rb, qb, dist=True, on_device=True)
        out.free()
        rb.free()
        qb.free()
    timed("closest-features --dist, 50M x 200M", ref.rows + qry.rows, closest)

    def closest_no():
        rb, qb = ref.load(kit, 3, COL_LINE), qry.load(kit, 3, COL_LINE)
        out = kit.closest(rb, qb, dist=True, no_overlaps=True, on_device=True)
        out.free()
        rb.free()
        qb.free()
    timed("closest-features --dist --no-overlaps, 50M x 200M", ref.rows + qry.rows, closest_no)
    del ref, qry
    kit.release_cached()
    torch.cuda.empty_cache()



def run_config5(kit, torch, timed, scale, SynthFile, MAP_SHAPE, REF_SHAPE, COL_LINE, COL_SCORE):
    # configuration 5: bedmap --mean over 1 B map rows on one GPU
    ref = SynthFile(kit, torch, int(10_000_000 * scale), 2, REF_SHAPE)
    mp = SynthFile(kit, torch, int(1_000_000_000 * scale), 1, MAP_SHAPE)

    def bedmap():
        rb, mb = ref.load(kit, 3, COL_LINE), mp.load(kit, 5, COL_SCORE)
        out = kit.bedmap(rb, mb, ["echo", "mean"], on_device=True)
        out.free()
        rb.free()
        mb.free()
    timed("bedmap --echo --mean, 10M x 1B", ref.rows + mp.rows, bedmap)


if __name__ == "__main__":
    main()
