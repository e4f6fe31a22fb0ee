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

N > 1 (torchrun): ONE dataset -- BASELINE.json configs[4], 10 M reference x 1 B map intervals -- cut by the repo's own
planner into N byte-balanced genomic ranges (cuts inside chromosomes, boundary halos; include/bedkit.h "range-sharded
bedmap"), one rank = one GPU = one range: strong scaling.  Rank 0 writes the text once into a shared-memory file, every
rank maps it and uploads / parses only its own slice (+ halos); the only collective is the allgather of N*N halo
coordinates (u64) per step.  Rank 0 checks the concatenated output against the unsharded run of the same dataset.

--impl reference: times the reference's own CPU implementation (oracle/_ref/bin/bedmap) on this box's host cores, on the
full 10 M x 100 M configuration (one process per chromosome, the reference's own scale-out).
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
    ap.add_argument("--big-map-rows", type=int, default=1_000_000_000, help="map rows of the N>1 dataset (configs[4])")
    ap.add_argument("--workload", default="auto", choices=["auto", "config2", "config5"],
                    help="auto: config2 at N=1, config5 (one dataset, range-sharded) at N>1")
    ap.add_argument("--cpu-sample-map-rows", type=int, default=100_000_000,
                    help="map rows of the CPU arm (default: the full configuration, same density)")
    ap.add_argument("--no-tool-e2e", action="store_true")
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
WORKLOAD2 = "bedmap --echo --count --mean --bases: 10M reference x 100M map intervals on 1 B200"
WORKLOAD5 = "bedmap --echo --count --mean --bases over a 1B-interval map set (10M reference rows) sharded by genomic range across %d B200"
SYNTH = os.path.join(ROOT, "bedops_b200", "bin", "synth-bed")


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


def shm_dir(prefix):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix=prefix, dir=base)


def write_cpu_arm_files(tmpdir, ref_rows, map_rows):
    """The CPU arm's input: the same synthetic shape at the same density, written by the host generator
    (bedops_b200/tools/synth-bed.cpp: no CUDA, the reference arm must not load the product's library)."""
    rp, mp = os.path.join(tmpdir, "ref.bed"), os.path.join(tmpdir, "map.bed")
    nref = int(subprocess.run([SYNTH, str(ref_rows), "2", "7.0", "1.0", "5", rp], capture_output=True, text=True, check=True).stdout)
    nmap = int(subprocess.run([SYNTH, str(map_rows), "1", "5.5", "1.0", "5", mp], capture_output=True, text=True, check=True).stdout)
    return rp, mp, nref, nmap


def big_first(chroms):
    """longest chromosomes first: the makespan of the per-chromosome processes is then close to total / cores"""
    from bedops_b200.synth import HG38
    return sorted(chroms, key=lambda c: -HG38[c])


def cpu_baseline(args):
    """The unmodified reference on this box's host cores, one full pass of the SAME configuration (10 M x 100 M, same
    density), per-chromosome parallel; plus a single-threaded pass over one chromosome for the per-core rate."""
    from bedops_b200.synth import HG38
    if not os.access(os.path.join(REFBIN, "bedmap"), os.X_OK) or not os.access(SYNTH, os.X_OK):
        return None
    td = shm_dir("bedkit_cpu_")
    try:
        rp, mp, nref, nmap = write_cpu_arm_files(td, args.cpu_sample_map_rows // 10, args.cpu_sample_map_rows)
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, len(HG38)))
        t = run_reference_bedmap(rp, mp, big_first(HG38), threads)
        out = {"value": (nref + nmap) / t, "unit": UNIT, "cores": threads, "kind": "reference",
               "sample": "one pass over %d ref x %d map rows (the whole configuration, same density); unmodified BEDOPS 2.4.26 "
                         "bedmap, one process per chromosome (--chrom), %d at a time, stdout to /dev/null" % (nref, nmap, threads)}
        # per-core rate: chr21 of the same files, one process
        c = "chr21"
        frac = HG38[c] / float(sum(HG38.values()))
        cmd = [os.path.join(REFBIN, "bedmap"), "--chrom", c, "--echo", "--count", "--mean", "--bases", rp, mp]
        t0 = time.perf_counter()
        with open(os.devnull, "wb") as dn:
            subprocess.run(cmd, stdout=dn, check=True)
        out["single_thread_value"] = (nref + nmap) * frac / (time.perf_counter() - t0)
        out["single_thread_sample"] = "%s of the same files, one process" % c
        return out
    finally:
        shutil.rmtree(td, ignore_errors=True)


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU path on this box's host cores, same metric and config: every
    step is one pass over the whole 10 M x 100 M configuration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bedops_b200.synth import HG38
    td = shm_dir("bedkit_ref_")
    try:
        rp, mp, nref, nmap = write_cpu_arm_files(td, args.cpu_sample_map_rows // 10, args.cpu_sample_map_rows)
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, len(HG38)))
        # bounded: the whole --steps/--warmup run has to end within a few minutes, a pass takes several seconds
        t1 = run_reference_bedmap(rp, mp, big_first(HG38), threads)            # first warm-up pass, also sizes the rest
        budget = 150.0
        warm = max(0, min(args.warmup - 1, int(budget * 0.2 / max(t1, 1e-3))))
        for _ in range(warm):
            run_reference_bedmap(rp, mp, big_first(HG38), threads)
        k = max(1, min(args.steps, int(budget * 0.8 / max(t1, 1e-3))))
        t = [run_reference_bedmap(rp, mp, big_first(HG38), threads) for _ in range(k)]
        total = sum(t)
        value = (nref + nmap) * k / total
        sample = ("each step = one pass over the whole configuration, %d ref x %d map rows at the benchmark's density; unmodified "
                  "BEDOPS 2.4.26 bedmap, one process per chromosome (--chrom), %d at a time; %d timed passes after %d warm-up "
                  "(bounded to ~%d s of CPU work)" % (nref, nmap, threads, k, warm + 1, int(budget)))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / k,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
            "config": {"workload": WORKLOAD2, "ref_rows": nref, "map_rows": nmap, "timed_passes": k},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
    finally:
        shutil.rmtree(td, ignore_errors=True)


def tool_e2e(kit, torch, ref_buf, ref_bytes, map_buf, map_bytes, units):
    """The real drop-in: wall clock of `bin/bedmap --echo --count --mean --bases ref.bed map.bed > /dev/null` (process start,
    CUDA context, file read, upload, kernels, download, write) on tmpfs files holding the benchmark's text, beside the same
    command line of the unmodified reference."""
    import numpy as np
    from bedops_b200._lib import tool_path
    td = shm_dir("bedkit_tool_")
    try:
        paths = []
        for name, buf, n in (("ref.bed", ref_buf, ref_bytes), ("map.bed", map_buf, map_bytes)):
            path = os.path.join(td, name)
            host = np.empty(n, dtype=np.uint8)
            kit.copy(host.ctypes.data, buf.data_ptr(), n)
            host.tofile(path)
            paths.append(path)
            del host
        cmd = ["--echo", "--count", "--mean", "--bases"] + paths
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            with open(os.devnull, "wb") as dn:
                subprocess.run([tool_path("bedmap")] + cmd, stdout=dn, check=True)
            t = time.perf_counter() - t0
            best = t if best is None else min(best, t)
        out = {"value": units / best, "unit": UNIT, "wall_s": best, "runs": 3,
               "command": "bedops_b200/bin/bedmap --echo --count --mean --bases ref.bed map.bed > /dev/null (files on tmpfs)"}
        if os.access(os.path.join(REFBIN, "bedmap"), os.X_OK):   # the same command line of the reference, one chromosome
            c = "chr21"
            from bedops_b200.synth import HG38
            frac = HG38[c] / float(sum(HG38.values()))
            t0 = time.perf_counter()
            with open(os.devnull, "wb") as dn:
                subprocess.run([os.path.join(REFBIN, "bedmap"), "--chrom", c] + cmd, stdout=dn, check=True)
            tr = time.perf_counter() - t0
            out["reference_single_process_value"] = units * frac / tr
            out["reference_sample"] = "reference bedmap --chrom %s on the same two files (%.1f s), scaled by the chromosome's share" % (c, tr)
        return out
    finally:
        shutil.rmtree(td, ignore_errors=True)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def ncu_traffic():
    """dram bytes of the k_parse launches of one step from the committed ncu capture (profiles/*_ncu_raw.csv of the
    current kernels: named in profiles/CURRENT); None when there is no capture for this build."""
    try:
        meta = json.load(open(os.path.join(ROOT, "profiles", "CURRENT.json")))
        return meta.get("k_parse_dram_bytes_per_step"), meta.get("source")
    except (OSError, ValueError):
        return None, None


# ---- the B200 arm -------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import bedops_b200
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
    workload = args.workload if args.workload != "auto" else ("config2" if world == 1 else "config5")
    try:
        if workload == "config2" and world == 1:
            run_config2(args, kit, torch, device, stream, numa)
        else:
            run_sharded(args, kit, torch, dist, device, stream, numa, world, rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


def kernel_split(kit, steps):
    names = ["k_count_rows", "k_scan_warps", "k_parse", "k_pmax_reduce", "k_pmax", "k_block_max", "k_map_group", "k_map_stats", "k_emit_len", "k_emit"]
    out = {}
    for k in names:
        ms, n = kit.profile_query(k)
        if n:
            out[k] = ms / steps
    return out


def run_config2(args, kit, torch, device, stream, numa):
    from bedops_b200._lib import COL_LINE, COL_SCORE
    from bedops_b200 import synth
    ref_buf, ref_bytes, nref = gen_device_bed(kit, torch, args.ref_rows, 2, *synth.REF_SHAPE, device)
    map_buf, map_bytes, nmap = gen_device_bed(kit, torch, args.map_rows, 1, *synth.MAP_SHAPE, device)
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

    sampler = ClockSampler(device.index)
    sampler.start()
    for _ in range(max(args.warmup, 3)):  # the first pass also fills the library's block cache
        out_bytes, out_rows = step_device()
    assert out_rows == nref, (out_rows, nref)
    kit.profile(True)
    torch.cuda.synchronize()
    sampler.mark_begin()
    l0 = kit.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    torch.cuda.synchronize()
    sampler.mark_end()
    clocks = sampler.stop()
    launches = kit.launches - l0
    ms = ev0.elapsed_time(ev1)
    split = kernel_split(kit, args.steps)
    kit.profile(False)
    value = units * args.steps / (ms * 1e-3)

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

        for _ in range(max(1, min(args.warmup, 3))):
            d2h = step_host()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = max(1, min(args.steps, 5))
        e0.record(stream)
        for _ in range(k):
            step_host()
        e1.record(stream)
        torch.cuda.synchronize()
        te = e0.elapsed_time(e1)
        e2e = {"value": units * k / (te * 1e-3), "unit": UNIT, "h2d_bytes_per_step": ref_bytes + map_bytes,
               "d2h_bytes_per_step": int(d2h), "steps": k, "ms_per_step": te / k}
        del ref_host, map_host

    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    parse_ms = split.get("k_parse", 0.0) * args.steps
    # k_parse per launch: the ref text and the map text are each parsed once per step
    alg_bytes_step = (map_bytes + 16 * nmap) + (ref_bytes + 16 * nref)   # text in + (start,end,score|line_off) out
    achieved = alg_bytes_step * args.steps / (parse_ms * 1e-3) / 1e9 if parse_ms > 0 else None
    traffic, traffic_src = (None, None)
    if args.ref_rows == 10_000_000 and args.map_rows == 100_000_000:
        traffic, traffic_src = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "k_parse", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "traffic_source": traffic_src, "algorithmic_bytes_per_step": alg_bytes_step, "kernel_ms_per_step": split,
                "whole_step_text_GBps": (ref_bytes + map_bytes + out_bytes) * args.steps / (ms * 1e-3) / 1e9,
                "whole_step_frac": (ref_bytes + map_bytes + out_bytes) * args.steps / (ms * 1e-3) / 1e9 / peak}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD2, "ref_rows": nref, "map_rows": nmap, "ref_text_bytes": ref_bytes,
                   "map_text_bytes": map_bytes, "out_text_bytes": int(out_bytes),
                   "l2": "inputs (%.2f GB text per step) are larger than the 126 MB L2; no explicit flush" % ((ref_bytes + map_bytes) / 1e9),
                   "numa_node": numa},
        "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_tool_e2e:
        try:
            line["tool_e2e"] = tool_e2e(kit, torch, ref_buf, ref_bytes, map_buf, map_bytes, units)
        except (OSError, subprocess.SubprocessError) as ex:
            line["tool_e2e"] = {"error": str(ex)[:200]}
    if not args.no_cpu_baseline:
        del ref_buf, map_buf
        cb = cpu_baseline(args)
        if cb:
            line["cpu_baseline"] = cb
    print(json.dumps(line))


# ---- N > 1: one dataset, range-sharded ----------------------------------------------------------------------------
def run_sharded(args, kit, torch, dist, device, stream, numa, world, rank):
    import ctypes
    import hashlib
    import mmap
    import numpy as np
    from bedops_b200._lib import COL_LINE, COL_SCORE
    from bedops_b200 import synth
    from bedops_b200.shard import make_plan, RangeShard, INF

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every rank builds the same dataset on its own GPU (same seeds): the device-resident mirror the `value` leg copies
    # from.  Rank 0 also writes it once into shared memory: the host text every rank maps (planner bisections, e2e source).
    ref_buf, ref_bytes, nref = gen_device_bed(kit, torch, args.ref_rows, 2, *synth.REF_SHAPE, device)
    map_buf, map_bytes, nmap = gen_device_bed(kit, torch, args.big_map_rows, 1, *synth.MAP_SHAPE, device)
    units = nref + nmap
    tag = os.environ.get("MASTER_PORT", "0")
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    paths = [os.path.join(shm, "bedkit_bench_%s_%s.bed" % (tag, n)) for n in ("ref", "map")]
    maps = []
    try:
        if rank == 0:
            for path, buf, n in zip(paths, (ref_buf, map_buf), (ref_bytes, map_bytes)):
                with open(path, "wb") as f:
                    f.truncate(n)
                with open(path, "r+b") as f:
                    mm = mmap.mmap(f.fileno(), n)
                    dst = np.frombuffer(mm, dtype=np.uint8)
                    step = 1 << 30
                    for off in range(0, n, step):
                        m = min(step, n - off)
                        kit.copy(dst.ctypes.data + off, buf.data_ptr() + off, m)
                    del dst
                    mm.flush()
                    mm.close()
        barrier()
        host = []
        for path, n in zip(paths, (ref_bytes, map_bytes)):
            f = open(path, "r+b")
            mm = mmap.mmap(f.fileno(), n)
            maps.append((f, mm))
            host.append(np.frombuffer(mm, dtype=np.uint8))
        ref_host, map_host = host
        plan = make_plan(ref_host, map_host, world)
        # pin this rank's own slices of the shared text (halos lie just outside: pinned with a margin)
        rt = torch.cuda.cudart()
        pinned = []
        page = 1 << 21
        for arr, off in ((ref_host, plan.ref_off), (map_host, plan.map_off)):
            lo = max(0, (off[rank] - (64 << 20)) // page * page)
            hi = min(arr.size, ((off[rank + 1] + (64 << 20)) + page - 1) // page * page)
            if hi > lo:
                err = rt.cudaHostRegister(arr.ctypes.data + lo, hi - lo, 0)
                if int(err) == 0:
                    pinned.append(arr.ctypes.data + lo)
                else:
                    print("[bench] rank %d: cudaHostRegister(%d bytes) failed: %s" % (rank, hi - lo, err), file=sys.stderr)

        gather_buf = [torch.empty(world, dtype=torch.int64, device=device) for _ in range(world)]

        def exchange(reach):
            # the path's only collective: N vectors of N halo coordinates (NCCL allgather over NVLink)
            mine = torch.tensor([min(v, 1 << 62) for v in reach], dtype=torch.int64, device=device)
            dist.all_gather(gather_buf, mine)
            rows = torch.stack(gather_buf).cpu().tolist()
            return [[INF if v >= (1 << 62) else v for v in r] for r in rows]

        phases = {"begin": 0.0, "exchange": 0.0, "finish": 0.0, "n": 0}

        def step(from_device, on_device, raw=False):
            t0 = time.perf_counter()
            sh = RangeShard(kit, plan, rank, ref_host, map_host, OPS, ref_fields=3, ref_cols=COL_LINE, map_fields=5,
                            map_cols=COL_SCORE, ref_src=(ref_buf.data_ptr(), ref_bytes) if from_device else None,
                            map_src=(map_buf.data_ptr(), map_bytes) if from_device else None, on_device=on_device)
            t1 = time.perf_counter()
            allr = exchange(sh.reach_list())
            t2 = time.perf_counter()
            out = sh.finish(allr, _raw=raw)
            t3 = time.perf_counter()
            phases["begin"] += t1 - t0; phases["exchange"] += t2 - t1; phases["finish"] += t3 - t2; phases["n"] += 1
            return sh, out

        def timed(from_device, on_device, steps, warmup):
            for _ in range(warmup):
                sh, out = step(from_device, on_device, raw=not on_device)
                if on_device:
                    out.free()
                else:
                    kit.free_text(out)
            barrier()
            if steps == 0:
                return 0.0, 0, 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            t0 = time.perf_counter()
            last = None
            for _ in range(steps):
                # e2e: the result text lands in the library's pinned host buffer (the tools write it to stdout from there)
                sh, out = step(from_device, on_device, raw=not on_device)
                last = (sh.bytes_in, out.nbytes if on_device else out.len)
                if on_device:
                    out.free()
                else:
                    kit.free_text(out)
            e1.record(stream)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            barrier()
            # the step mixes host work (bisections, the exchange) with device work: the device-event time and the host wall
            # clock of the same region agree; the larger one, max over ranks, is reported
            t = torch.tensor([max(e0.elapsed_time(e1), wall)], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            b = torch.tensor([last[0], last[1]], device=device, dtype=torch.int64)
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            return float(t.item()), int(b[0].item()), int(b[1].item())

        sampler = ClockSampler(device.index)
        sampler.start()
        timed(True, True, 0, max(args.warmup, 3))
        kit.profile(True)
        l0 = kit.launches
        sampler.mark_begin()
        phases.update(begin=0.0, exchange=0.0, finish=0.0, n=0)
        ms, bytes_in, out_bytes = timed(True, True, args.steps, 0)
        host_phases_ms = {k: 1e3 * phases[k] / max(1, phases["n"]) for k in ("begin", "exchange", "finish")}
        sampler.mark_end()
        clocks = sampler.stop()
        launches = kit.launches - l0
        split = kernel_split(kit, args.steps)
        kit.profile(False)
        value = units * args.steps / (ms * 1e-3)

        e2e = None
        if not args.no_e2e:
            k = max(1, min(args.steps, 5))
            phases.update(begin=0.0, exchange=0.0, finish=0.0, n=0)
            te, h2d, d2h = timed(False, False, k, max(1, min(args.warmup, 2)))
            e2e_phases_ms = {k_: 1e3 * phases[k_] / max(1, phases["n"]) for k_ in ("begin", "exchange", "finish")}
            e2e = {"value": units * k / (te * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "steps": k, "ms_per_step": te / k, "rank0_host_phases_ms": e2e_phases_ms,
                   "note": "bytes summed over ranks: every rank uploads only its own slice of the shared text (+ halos)"}

        # ---- the sharded output is the unsharded output: rank 0 maps the whole dataset alone and compares part by part
        sh, part = step(False, False)
        digest = torch.tensor(list(hashlib.sha256(part).digest()) + [0] * 0, dtype=torch.uint8, device=device)
        size = torch.tensor([len(part)], dtype=torch.int64, device=device)
        sizes = [torch.empty(1, dtype=torch.int64, device=device) for _ in range(world)]
        digests = [torch.empty(32, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.all_gather(sizes, size)
        dist.all_gather(digests, digest)
        verified = None
        single = None
        if rank == 0:
            sizes = [int(x.item()) for x in sizes]
            ref = kit.load_device(ref_buf.data_ptr(), ref_bytes, 3, COL_LINE)
            mp = kit.load_device(map_buf.data_ptr(), map_bytes, 5, COL_SCORE)
            whole = kit.bedmap(ref, mp, OPS)
            ok, off = len(whole) == sum(sizes), 0
            for n, d in zip(sizes, digests):
                ok = ok and hashlib.sha256(whole[off:off + n]).digest() == bytes(d.cpu().tolist())
                off += n
            verified = bool(ok)
            del whole
            # the same dataset on ONE GPU, device-resident: the strong-scaling denominator
            def one():
                r_ = kit.load_device(ref_buf.data_ptr(), ref_bytes, 3, COL_LINE)
                m_ = kit.load_device(map_buf.data_ptr(), map_bytes, 5, COL_SCORE)
                o_ = kit.bedmap(r_, m_, OPS, on_device=True)
                o_.free(); r_.free(); m_.free()
            ref.free(); mp.free()
            one()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(3):
                one()
            a1.record(stream)
            torch.cuda.synchronize()
            single = {"value": units * 3 / (a0.elapsed_time(a1) * 1e-3), "unit": UNIT, "ms_per_step": a0.elapsed_time(a1) / 3,
                      "note": "the same 10M x 1B dataset unsharded on one GPU, device-resident (the N=1 point of this workload)"}
        barrier()
        for pa in pinned:
            rt.cudaHostUnregister(pa)
        if rank != 0:
            return
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        parse_ms = split.get("k_parse", 0.0)
        # rank 0's k_parse launches of one step parse its slices: text in + 16 B per row out
        r0_bytes = (plan.map_off[1] - plan.map_off[0]) + (plan.ref_off[1] - plan.ref_off[0])
        r0_rows = units / world
        alg = r0_bytes + 16 * r0_rows
        achieved = alg / (parse_ms * 1e-3) / 1e9 if parse_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32/f64", "data": "synthetic",
            "config": {"workload": WORKLOAD5 % world, "ref_rows": nref, "map_rows": nmap, "ref_text_bytes": ref_bytes,
                       "map_text_bytes": map_bytes, "out_text_bytes": out_bytes,
                       "sharding": "ONE dataset; bk_shard_plan_make cuts it into %d byte-balanced genomic ranges (cuts inside "
                                   "chromosomes), left/right halos by the prefix-max-end index; per step one NCCL allgather "
                                   "of %d u64; outputs in rank order" % (world, world * world),
                       "cuts": [{"chrom": plan.cuts[k].chrom.decode(), "start": int(plan.cuts[k].coord)} for k in range(world - 1)],
                       "sharded_output_equals_unsharded": verified,
                       "l2": "every rank's slice (%.1f GB) is larger than the 126 MB L2; no explicit flush" % (map_bytes / world / 1e9),
                       "numa_node": numa},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_parse", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                         "note": "rank 0's launches", "kernel_ms_per_step": split},
            "same_workload_one_gpu": single, "rank0_host_phases_ms": host_phases_ms,
        }
        if e2e:
            line["e2e"] = e2e
        print(json.dumps(line))
    finally:
        for f, mm in maps:
            try:
                mm.close()
            except (BufferError, ValueError):
                pass
            f.close()
        if world > 1:
            try:
                dist.barrier()
            except Exception:
                pass
        if rank == 0:
            for path in paths:
                try:
                    os.unlink(path)
                except OSError:
                    pass


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
