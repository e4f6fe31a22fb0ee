#!/usr/bin/env python
"""bench.py -- bedmap --echo --count --mean --bases throughput (intervals/s) on B200, with roofline and CPU baseline.

Workload (BASELINE.json configs[1]): 10 M reference x 100 M map intervals, synthetic sorted BED5 of hg38 shape
(SURVEY.md 8d), generated ON DEVICE by torch RNG + the library's own BED writer.  One "step" = one full pass of
the hot path: parse the reference text, parse the map text, build the prefix-max index, reduce every reference
row's candidate window, emit the output text.

  value   whole-job (ref+map) intervals/s with the input text resident in HBM (device pointers in, device text out)
  e2e     the same step through the host-buffer C ABI (bk_load_bed / bk_bedmap): pinned host text in, H2D inside
          the timed region, result text copied back to pinned host memory inside the timed region
  roofline  the dominant kernel (k_parse over the map text): algorithmic bytes = text bytes + SoA bytes written,
          divided by its live CUDA-event duration (events recorded by the library on its launching stream)
  cpu_baseline  the UNMODIFIED reference bedmap (oracle/_ref/bin) on a bounded sample of the same workload, on the
          host cores of this box (per-chromosome parallel with --chrom, the reference's own scale-out mechanism)

N > 1 (torchrun): every rank owns one GPU and an independent genomic shard of the same shape (weak scaling, no
data-path collective -- the path shards by genomic range); timing is barrier + max over ranks.

--impl reference: times the reference's own CPU implementation (oracle/_ref/bin/bedmap) on this box's host cores.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")

OPS = ["echo", "count", "mean", "bases"]
METRIC = "bedmap --echo --count --mean --bases (ref+map) intervals/sec"
UNIT = "intervals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-rows", type=int, default=10_000_000)
    ap.add_argument("--map-rows", type=int, default=100_000_000)
    ap.add_argument("--cpu-sample-map-rows", type=int, default=6_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---- synthetic workload on the device -----------------------------------------------------------------------
def gen_device_bed(kit, torch, n_total, seed, mu, sigma, device):
    """Sorted BED5 text in HBM (one uint8 tensor) of ~n_total rows; same distributions as bedops_b200.synth."""
    from bedops_b200.synth import HG38
    total = float(sum(HG38.values()))
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    parts, k = [], 0
    for c in sorted(HG38):
        size = HG38[c]
        m = int(round(n_total * size / total))
        if m == 0:
            continue
        length = torch.empty(m, device=device, dtype=torch.float32).log_normal_(mu, sigma, generator=g).to(torch.int64).clamp_(min=1)
        s = torch.randint(0, size - 1, (m,), device=device, dtype=torch.int64, generator=g)
        e = torch.minimum(s + length, torch.tensor(size, device=device))
        e = torch.where(e <= s, s + 1, e)
        key, _ = torch.sort(s * (1 << 32) + e)
        s32 = (key >> 32).to(torch.int32)
        e32 = (key & 0xFFFFFFFF).to(torch.int32)
        sc = torch.randint(0, 1000, (m,), device=device, dtype=torch.int32, generator=g)
        torch.cuda.synchronize()
        t = kit.format_bed_device(c.encode(), s32.data_ptr(), e32.data_ptr(), sc.data_ptr(), m, k)
        parts.append(t)
        k += m
        del length, s, e, key, s32, e32, sc
    nbytes = sum(p.nbytes for p in parts)
    buf = torch.empty(nbytes + 64, dtype=torch.uint8, device=device)
    off = 0
    for p in parts:
        kit.copy(buf.data_ptr() + off, p.ptr, p.nbytes)
        off += p.nbytes
        p.free()
    return buf, nbytes, k


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the e2e leg are
    placed there (eight ranks sharing one node's memory controller cost the N=8 e2e a third of its rate).  Best effort:
    returns the node number or None."""
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                    # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except (OSError, ValueError, subprocess.SubprocessError):
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []   # lines: (monotonic time, text)
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.monotonic()

    def mark_end(self):
        self.t1 = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        # nvidia-smi needs 0.1-0.5 s to produce its first line (longer on an 8-GPU box), so it is started before the
        # warm-up; only the samples taken between mark_begin() and mark_end() -- the timed region -- are used
        lines = [l for t, l in self.lines if self.t0 is None or (self.t0 <= t <= (self.t1 or t))]
        if not lines and self.lines:   # region shorter than the polling period: the sample closest to it
            mid = ((self.t0 or 0) + (self.t1 or 0)) / 2
            lines = [min(self.lines, key=lambda tl: abs(tl[0] - mid))[1]]
        sm, mx, reasons = [], None, set()
        for l in lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference on the host cores --------------------------------------------------------------------------
def run_reference_bedmap(ref_path, map_path, chroms, threads):
    """One pass of the reference bedmap over the files, one process per chromosome (--chrom), `threads` at a time.
    Returns wall seconds."""
    cmd = [os.path.join(REFBIN, "bedmap"), "--echo", "--count", "--mean", "--bases"]
    t0 = time.perf_counter()
    if threads <= 1:
        with open(os.devnull, "wb") as dn:
            subprocess.run(cmd + [ref_path, map_path], stdout=dn, check=True)
        return time.perf_counter() - t0
    pending = list(chroms)
    running = []
    dn = open(os.devnull, "wb")
    while pending or running:
        while pending and len(running) < threads:
            c = pending.pop(0)
            running.append(subprocess.Popen(cmd[:1] + ["--chrom", c] + cmd[1:] + [ref_path, map_path], stdout=dn))
        running = [p for p in running if p.poll() is None]
        time.sleep(0.002)
    dn.close()
    return time.perf_counter() - t0


def write_sample_files(tmpdir, ref_rows, map_rows):
    """Bounded sample of the workload for the CPU arm: same generator (numpy leg), written to a tmpfs if possible."""
    from bedops_b200 import synth
    rp, mp = os.path.join(tmpdir, "ref.bed"), os.path.join(tmpdir, "map.bed")
    with open(rp, "wb") as f:
        f.write(synth.bed_text(ref_rows, 2, synth.REF_SHAPE))
    with open(mp, "wb") as f:
        f.write(synth.bed_text(map_rows, 1, synth.MAP_SHAPE))
    nref = sum(1 for _ in open(rp, "rb"))
    nmap = sum(1 for _ in open(mp, "rb"))
    return rp, mp, nref, nmap


def cpu_baseline(sample_map_rows, steps=1):
    from bedops_b200.synth import HG38
    if not os.access(os.path.join(REFBIN, "bedmap"), os.X_OK):
        return None
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    td = tempfile.mkdtemp(prefix="bedkit_cpu_", dir=base)
    try:
        rp, mp, nref, nmap = write_sample_files(td, sample_map_rows // 10, sample_map_rows)
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, len(HG38)))
        best = min(run_reference_bedmap(rp, mp, sorted(HG38), threads) for _ in range(steps))
        single = run_reference_bedmap(rp, mp, sorted(HG38), 1) if sample_map_rows <= 8_000_000 else None
        out = {"value": (nref + nmap) / best, "unit": UNIT, "cores": threads, "kind": "reference",
               "sample": "%d ref x %d map rows of the same synthetic shape; unmodified BEDOPS 2.4.26 bedmap, one process "
                         "per chromosome (--chrom), %d at a time, stdout to /dev/null" % (nref, nmap, threads)}
        if single:
            out["single_thread_value"] = (nref + nmap) / single
        return out
    finally:
        shutil.rmtree(td, ignore_errors=True)


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU path on this box's host cores, same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bedops_b200.synth import HG38
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    td = tempfile.mkdtemp(prefix="bedkit_ref_", dir=base)
    try:
        rp, mp, nref, nmap = write_sample_files(td, args.cpu_sample_map_rows // 10, args.cpu_sample_map_rows)
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, len(HG38)))
        for _ in range(min(args.warmup, 1)):
            run_reference_bedmap(rp, mp, sorted(HG38), threads)
        t = [run_reference_bedmap(rp, mp, sorted(HG38), threads) for _ in range(args.steps)]
        total = sum(t)
        value = (nref + nmap) * args.steps / total
        sample = ("each step = %d ref x %d map rows (bounded sample of the 10M x 100M workload); unmodified BEDOPS 2.4.26 "
                  "bedmap, one process per chromosome (--chrom), %d at a time" % (nref, nmap, threads))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
            "config": {"workload": "bedmap --echo --count --mean --bases: 10M reference x 100M map intervals on 1 B200",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
    finally:
        shutil.rmtree(td, ignore_errors=True)


# ---- the B200 arm -------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import bedops_b200
    from bedops_b200._lib import COL_LINE, COL_SCORE
    from bedops_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)      # host buffers of the e2e leg: first touch on the GPU's own node
    kit = bedops_b200.BedKit(local)          # raises without a B200: there is no CPU fallback
    stream = torch.cuda.current_stream(device)
    kit.set_stream(stream.cuda_stream)

    # one shard per rank: same shape, different seed (weak scaling over genomic shards)
    ref_buf, ref_bytes, nref = gen_device_bed(kit, torch, args.ref_rows, 2 + 100 * rank, *synth.REF_SHAPE, device)
    map_buf, map_bytes, nmap = gen_device_bed(kit, torch, args.map_rows, 1 + 100 * rank, *synth.MAP_SHAPE, device)
    units = nref + nmap

    def step_device():
        ref = kit.load_device(ref_buf.data_ptr(), ref_bytes, 3, COL_LINE)
        mp = kit.load_device(map_buf.data_ptr(), map_bytes, 5, COL_SCORE)
        out = kit.bedmap(ref, mp, OPS, on_device=True)
        nb, rows = out.nbytes, out.rows
        out.free()
        ref.free()
        mp.free()
        return nb, rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 1)):  # the first pass also fills the library's block cache
        out_bytes, out_rows = step_device()
    assert out_rows == nref, (out_rows, nref)
    kit.profile(True)
    barrier()
    sampler.mark_begin()
    l0 = kit.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    launches = kit.launches - l0
    ms = ev0.elapsed_time(ev1)
    parse_ms, parse_n = kit.profile_query("k_parse")
    stats_ms, stats_n = kit.profile_query("k_map_stats")
    emit_ms, emit_n = kit.profile_query("k_emit")
    emit_len_ms, _ = kit.profile_query("k_emit_len")
    pmax_ms, pmax_n = kit.profile_query("k_pmax")
    pmax_ms += kit.profile_query("k_pmax_reduce")[0]
    count_ms, _ = kit.profile_query("k_count_rows")
    scan_ms, _ = kit.profile_query("k_scan_warps")
    kit.profile(False)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * units * args.steps / (ms_max * 1e-3)

    # ---- end to end through the host-buffer ABI ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        ref_host = torch.empty(ref_bytes, dtype=torch.uint8, pin_memory=True)
        map_host = torch.empty(map_bytes, dtype=torch.uint8, pin_memory=True)
        ref_host.copy_(ref_buf[:ref_bytes])
        map_host.copy_(map_buf[:map_bytes])
        torch.cuda.synchronize()

        def step_host():
            # the reference-facing call with HOST buffers: bk_bedmap_host uploads, parses, maps and downloads chromosome
            # group by chromosome group (H2D, kernels and D2H overlapped); result text lands in pinned host memory
            text = kit.bedmap_host(ref_host.data_ptr(), ref_bytes, 3, COL_LINE, map_host.data_ptr(), map_bytes, 5,
                                   COL_SCORE, OPS, _raw=True)
            n = text.len
            kit.free_text(text)
            return n

        for _ in range(max(1, min(args.warmup, 2))):
            d2h = step_host()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = max(1, min(args.steps, 3))
        e0.record(stream)
        for _ in range(k):
            step_host()
        e1.record(stream)
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * units * k / (float(te.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": ref_bytes + map_bytes, "d2h_bytes_per_step": int(d2h), "steps": k,
               "ms_per_step": float(te.item()) / k}
        del ref_host, map_host

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # k_parse per launch: the ref text and the map text are each parsed once per step
    alg_bytes_step = (map_bytes + 16 * nmap) + (ref_bytes + 16 * nref)   # text in + (start,end,score|line_off) out
    achieved = alg_bytes_step * args.steps / (parse_ms * 1e-3) / 1e9 if parse_ms > 0 else None
    # dram__bytes_read.sum + dram__bytes_write.sum of the two k_parse launches of one step (reference file + map file),
    # from the ncu --set full capture of this very command (profiles/r01_final_ncu_raw.csv); only valid for the default workload
    traffic = None
    if args.ref_rows == 10_000_000 and args.map_rows == 100_000_000:
        traffic = int((0.377873 + 0.143716 + 3.870698 + 1.609264) * 1e9)
    roofline = {"bound": "hbm", "kernel": "k_parse", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "traffic_source": "ncu dram bytes, both k_parse launches of a step (profiles/r01_final_ncu_raw.csv)" if traffic else None,
                "algorithmic_bytes_per_step": alg_bytes_step,
                "kernel_ms_per_step": {"k_count_rows": count_ms / args.steps, "k_scan_warps": scan_ms / args.steps,
                                       "k_parse": parse_ms / args.steps, "k_pmax_reduce+k_pmax": pmax_ms / args.steps,
                                       "k_map_stats": stats_ms / args.steps, "k_emit_len": emit_len_ms / args.steps, "k_emit": emit_ms / args.steps},
                "whole_step_text_GBps": (ref_bytes + map_bytes + out_bytes) * args.steps / (ms_max * 1e-3) / 1e9}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": "bedmap --echo --count --mean --bases: 10M reference x 100M map intervals on 1 B200",
                   "ref_rows": nref, "map_rows": nmap, "ref_text_bytes": ref_bytes, "map_text_bytes": map_bytes,
                   "out_text_bytes": int(out_bytes), "per_gpu": True,
                   "l2": "inputs (%.2f GB text per step) are larger than the 126 MB L2; no explicit flush" % ((ref_bytes + map_bytes) / 1e9),
                   "sharding": "one independent genomic shard of this shape per GPU, no data-path collective",
                   "numa_node": numa},
        "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(args.cpu_sample_map_rows)
        if cb:
            line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    try:
        main()
    finally:
        _real_stdout.flush()
