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


# ---- cuts inside chromosomes with boundary halos (bk_shard_plan_make / bk_find_start; protocol: include/bedkit.h) -------
def _starts(text):
    out, pos = [], 0
    for line in text.split(b"\n")[:-1]:
        f = line.split()
        if f:
            out.append((pos, f[0], int(f[1])))
        pos += len(line) + 1
    return out


def test_find_start_matches_a_linear_scan(synth_files):
    from bedops_b200.shard import chrom_index, find_start
    import numpy as np
    rng = np.random.default_rng(7)
    for name in ("m.bed", "r.bed"):
        text = synth_files[name]
        recs = _starts(text)
        for chrom, b, e in chrom_index(text)[:6]:
            mine = [(p, s) for p, c, s in recs if c == chrom]
            probes = [0, 1, mine[0][1], mine[-1][1], mine[-1][1] + 1, 1 << 40] + [int(x) for x in rng.integers(0, mine[-1][1] + 5, 40)]
            probes += [s for _, s in mine[:: max(1, len(mine) // 25)]]
            for x in probes:
                exp = next((p for p, s in mine if s >= x), e)
                assert find_start(text, b, e, x) == exp, (name, chrom, x)
    # blank lines, equal starts, a range that is one line
    t = b"c\t5\t6\n\n\nc\t5\t9\nc\t7\t8\n  \nc\t7\t9\nc\t20\t21\n"
    for x, exp in ((0, 0), (5, 0), (6, 14), (7, 14), (8, 29), (20, 29), (21, len(t))):
        assert find_start(t, 0, len(t), x) == exp, x
    assert find_start(t, 14, 20, 7) == 14 and find_start(t, 14, 20, 8) == 20


def test_range_plan_cuts_every_file_at_the_same_genomic_positions(synth_files):
    from bedops_b200.shard import make_plan
    ref, mp_ = synth_files["r.bed"], synth_files["m.bed"]
    for n in (1, 2, 3, 5, 8, 16, 64):
        plan = make_plan(ref, mp_, n)
        assert plan.n_shards == n
        for text, off in ((ref, plan.ref_off), (mp_, plan.map_off)):
            offs = [off[k] for k in range(n + 1)]
            assert offs == sorted(offs) and offs[0] == 0 and offs[n] == len(text)
            recs = _starts(text)
            for k in range(n - 1):
                cut = plan.cuts[k]
                key = (cut.chrom, cut.coord)
                if cut.at_end:
                    assert offs[k + 1] == len(text)
                    continue
                assert all(((c, s) >= key) == (p >= offs[k + 1]) for p, c, s in recs), (n, k)
        sizes = [plan.map_off[k + 1] - plan.map_off[k] for k in range(n)]
        if n <= 16:   # the map file is the larger one: its bytes are what the cuts balance
            assert max(sizes) <= len(mp_) / n + 4096, (n, sizes)


def range_shard_with_oracle(rank, world, ref, mp_, argv_ops, overlap, gather):
    """The range-sharding protocol of pipeline.cu (bk_bedmap_shard_begin / _finish) with the oracle standing in for the
    device: the cuts, byte offsets and bisections are the product's host planner; reach / max end are read off the
    oracle-parsed rows.  gather(list) -> list of every rank's list."""
    from bedops_b200.shard import make_plan, find_start, INF
    plan = make_plan(ref, mp_, world)
    pad = overlap[1] if overlap[0] == "range" else 0
    r_own = ref[plan.ref_off[rank]:plan.ref_off[rank + 1]]
    m0, m1 = plan.map_off[rank], plan.map_off[rank + 1]
    if rank + 1 < world and not plan.cuts[rank].at_end and plan.cuts[rank].coord > 0:
        c = plan.cuts[rank]
        ends = [r.end for r in O.parse_bed(r_own, 3) if r.chrom == c.chrom]
        if ends and max(ends) + pad > c.coord:
            m1 = find_start(mp_, m1, plan.map_chrom_end[rank], max(ends) + pad)
    m_main = mp_[m0:m1]
    rows = O.parse_bed(m_main, 3)
    reach = [INF] * world
    for j in range(rank + 1, world):
        c = plan.cuts[j - 1]
        if c.at_end or c.coord == 0:
            continue
        pos = max(0, c.coord - pad)
        hit = [r.start for r in rows if r.chrom == c.chrom and r.end > pos]
        if hit and min(hit) < c.coord:
            reach[j] = min(hit)
    allr = gather(reach)
    if rank > 0:
        c = plan.cuts[rank - 1]
        s = min(allr[i][rank] for i in range(rank))
        if not c.at_end and c.coord > 0 and s < c.coord:
            h0 = find_start(mp_, plan.map_chrom_begin[rank - 1], m0, s)
            m_main = mp_[h0:m0] + m_main
    return O.bedmap(r_own, m_main, argv_ops, overlap=overlap)


def _range_worker(rank, world, port, cases, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bed_oracle as OO
    from test_shard import range_shard_with_oracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)

    def gather(v):
        out = [None] * world
        dist.all_gather_object(out, v)
        return out
    bad = []
    for k, (ref, mp_, ops, overlap) in enumerate(cases):
        part = range_shard_with_oracle(rank, world, ref, mp_, ops, overlap, gather)
        parts = gather(part)
        if rank == 0 and b"".join(parts) != OO.bedmap(ref, mp_, ops, overlap=overlap):
            bad.append(k)
    if rank == 0:
        q.put(bad)
    dist.barrier()
    dist.destroy_process_group()


def halo_cases(synth_files):
    import numpy as np
    cases = [(synth_files["r.bed"], synth_files["m.bed"], ["echo", "count", "mean", "bases"], ("bp", 1)),
             (synth_files["r.bed"], synth_files["m.bed"], ["count", "echo-map-id"], ("range", 2000))]
    # one chromosome only (the whole-chromosome planner could not split it at all); a nested interval that spans every
    # cut, long reference rows that reach far behind their shard's right cut, duplicate starts around the cuts
    rng = np.random.default_rng(11)
    s = np.sort(rng.integers(0, 200000, 4000))
    e = s + np.maximum(1, rng.lognormal(4.0, 1.5, 4000).astype(np.int64))
    rows = [(0, 199999)] + list(zip(s.tolist(), e.tolist())) + [(100000, 100001)] * 30
    rows.sort()
    mp_ = "".join("chrZ\t%d\t%d\tid%d\t%d\n" % (a, b, k, k % 97) for k, (a, b) in enumerate(rows)).encode()
    rs = np.sort(rng.integers(0, 200000, 900))
    re_ = rs + np.maximum(1, rng.lognormal(6.0, 1.8, 900).astype(np.int64))
    ref = "".join("chrZ\t%d\t%d\n" % (a, b) for a, b in zip(rs.tolist(), re_.tolist())).encode()
    cases.append((ref, mp_, ["echo", "count", "sum", "bases", "echo-map-id"], ("bp", 1)))
    cases.append((ref, mp_, ["count", "bases"], ("range", 500)))
    cases.append((ref, mp_, ["count"], ("fraction-map", 0.5)))
    # the reference file is the larger one: the cuts come from it
    cases.append((mp_, ref, ["count", "bases"], ("bp", 1)))
    return cases


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_range_shards_with_halos_concatenate_to_the_unsharded_result(synth_files, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_range_worker, args=(r, world, port, halo_cases(synth_files), q)) for r in range(world)]
    for p in procs:
        p.start()
    bad = q.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
    assert bad == []


@pytest.mark.parametrize("world", [8, 13])
def test_many_range_shards_in_one_process_concatenate_to_the_unsharded_result(synth_files, world):
    """The same protocol at the shard counts `bench.py --gpus 8` uses (and one that divides nothing), ranks as threads
    and the allgather as a barrier over a shared table: halos that begin more than one shard back, shards without reference
    rows, cuts in a single chromosome."""
    import threading
    barrier = threading.Barrier(world)
    cases = halo_cases(synth_files)
    for ref, mp_, ops, overlap in (cases if world == 8 else cases[2:]):   # 13 shards: the single-chromosome cases only (time)
        table, parts, errors = [None] * world, [None] * world, []

        def rank_main(rank):
            def gather(v):
                table[rank] = v
                barrier.wait(timeout=300)
                snapshot = list(table)
                barrier.wait(timeout=300)
                return snapshot
            try:
                parts[rank] = range_shard_with_oracle(rank, world, ref, mp_, ops, overlap, gather)
            except Exception as e:   # a dead rank must not leave the others at the barrier
                errors.append((rank, repr(e)))
                barrier.abort()

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=600)
        assert not errors, errors
        assert b"".join(parts) == O.bedmap(ref, mp_, ops, overlap=overlap), (world, ops, overlap)
