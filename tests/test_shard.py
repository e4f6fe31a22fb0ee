"""CPU tests of the multi-GPU host logic: chromosome index, balanced contiguous plan, and -- with two
gloo ranks, each computing its own shard with the oracle standing in for the device -- that the outputs concatenated
in shard order are byte-identical to the unsharded run for every tool on the path."""
import os
import sys

import pytest

from conftest import ROOT
import bed_oracle as O
import oracle_cli


def test_chrom_index_matches_a_linear_scan(synth_files):
    from bedops_b200.shard import chrom_index
    for name in ("m.bed", "r.bed", "m3.bed"):
        text = synth_files[name]
        got = chrom_index(text)
        exp, pos = [], 0
        for line in text.split(b"\n")[:-1]:
            c = line.split(b"\t")[0]
            if not exp or exp[-1][0] != c:
                if exp:
                    exp[-1][2] = pos
                exp.append([c, pos, None])
            pos += len(line) + 1
        exp[-1][2] = pos
        assert got == [tuple(e) for e in exp]
    # blank lines, an unterminated tail and a single chromosome
    t = b"\n\nchrA\t1\t2\nchrA\t3\t4\n\nchrB\t1\t2\nchrC\t5\t6"
    assert chrom_index(t) == [(b"chrA", 2, 21), (b"chrB", 21, 30)]
    assert chrom_index(b"") == []


def test_plan_is_contiguous_balanced_and_complete(synth_files):
    from bedops_b200.shard import chrom_index, plan_shards
    files = [synth_files["r.bed"], synth_files["m.bed"]]
    total = sum(len(f) for f in files)
    for n in (1, 2, 3, 4, 8, 24, 40):
        shards = plan_shards(files, n)
        assert len(shards) == n
        names = [c for s in shards for c in s["chroms"]]
        assert names == sorted({c for f in files for c, _, _ in chrom_index(f)})
        for k, f in enumerate(files):
            assert b"".join(f[s["slices"][k][0]:s["slices"][k][1]] for s in shards) == f
        biggest = max(s["load"] for s in shards)
        assert biggest >= total / n
        if n <= 8:
            assert biggest <= 1.6 * total / n  # hg38: chr1 is 8 % of the genome


def _worker(rank, world, port, cases, files, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_cli as oc
    from bedops_b200.shard import plan_shards
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ok = True
    for tool, argv in cases:
        names = [a for a in argv if a in files]
        shards = plan_shards([files[n] for n in names], world)
        mine = shards[rank]
        local = {n: files[n][mine["slices"][k][0]:mine["slices"][k][1]] for k, n in enumerate(names)}
        part = oc.run(tool, argv, local)
        parts = [None] * world
        dist.all_gather_object(parts, part)
        if rank == 0:
            ok = ok and b"".join(parts) == oc.run(tool, argv, files)
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_outputs_concatenate_to_the_unsharded_result(synth_files):
    import torch.multiprocessing as mp
    cases = [("bedmap", ["--echo", "--count", "--mean", "--bases", "r.bed", "m.bed"]),
             ("bedmap", ["--count", "--echo-map-id", "r.bed", "u.bed"]),
             ("bedops", ["-m", "m.bed", "m2.bed"]), ("bedops", ["-i", "r.bed", "m.bed"]),
             ("bedops", ["-e", "50%", "r.bed", "m.bed"]), ("bedops", ["-n", "1", "r.bed", "m3.bed"]),
             ("closest-features", ["--dist", "r.bed", "m.bed"])]
    files = {k: v for k, v in synth_files.items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cases, files, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert ok
