#!/usr/bin/env python3
"""Host-side timeline of one bedops call at scale (BEDKIT_TRACE=1 adds the library's allocation costs on stderr)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bedops_b200
from bedops_b200._lib import COL_LINE
from test_gpu_scale import SynthFile, MAP_SHAPE
n, nfiles = int(float(sys.argv[1])), int(sys.argv[2])
kit = bedops_b200.BedKit(0)
files = [SynthFile(kit, torch, n, s, MAP_SHAPE) for s in (1, 3, 4, 5)[:nfiles]]
def sync(): torch.cuda.synchronize(); kit.sync()
for op, thr, pct in (("element-of", 1, False), ("not-element-of", 1.0, True)):
    for rep in range(3):
        sync(); t0 = time.perf_counter()
        beds = [f.load(kit, 3, COL_LINE if k == 0 else 0) for k, f in enumerate(files)]
        sync(); t1 = time.perf_counter()
        print("-- setop", op, file=sys.stderr, flush=True)
        out = kit.setop(op, beds, thr, pct, on_device=True)
        sync(); t2 = time.perf_counter()
        nb = out.nbytes
        out.free()
        for b in beds: b.free()
        sync(); t3 = time.perf_counter()
        print("%s rep %d: load %.1f ms  setop %.1f ms  free %.1f ms  out %.2f GB" % (op, rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, nb/1e9), file=sys.stderr, flush=True)
