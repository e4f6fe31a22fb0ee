import sys, os
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, numpy as np
import bedops_b200
from test_gpu_scale import SynthFile, MAP_SHAPE, device_text_to_tensor
kit = bedops_b200.BedKit(0)
f = SynthFile(kit, torch, int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000, 1, MAP_SHAPE)
out = kit.sort_bed_device(f.buf.data_ptr(), f.nbytes, on_device=True)
print("in", f.nbytes, f.rows, "out", out.nbytes, out.rows)
a = f.buf[:f.nbytes].cpu().numpy().tobytes()
b = device_text_to_tensor(kit, torch, out).cpu().numpy().tobytes()
la, lb = a.split(b"\n"), b.split(b"\n")
print(len(la), len(lb))
from collections import Counter
ca, cb = Counter(la), Counter(lb)
missing = list((ca - cb).items())[:10]
extra = list((cb - ca).items())[:10]
print("missing", len(ca - cb), missing)
print("extra", len(cb - ca), extra)
n = 0
for i, (x, y) in enumerate(zip(la, lb)):
    if x != y:
        print(i, x, y, la[i-1:i+3], lb[i-1:i+3])
        n += 1
        if n > 5: break
