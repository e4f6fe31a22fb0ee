import sys, os
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, numpy as np
import bedops_b200
from test_gpu_scale import SynthFile, MAP_SHAPE, device_text_to_tensor
kit = bedops_b200.BedKit(0)
n = int(sys.argv[1])
f = SynthFile(kit, torch, n, 1, MAP_SHAPE)
blocks = []
for k, ch in enumerate(reversed(f.chroms)):
    b0, b1 = ch["b0"], ch["b1"]
    if ch["name"] == "chr1" and len(sys.argv) > 2:
        mid = (b0 + b1) // 2
        window = f.buf[mid:mid + 4096].cpu().numpy().tobytes()
        cut = mid + window.index(b"\n") + 1
        blocks += [(cut, b1), (b0, cut)]
    else:
        blocks.append((b0, b1))
shuffled = torch.empty(f.nbytes + 64, dtype=torch.uint8, device="cuda:0")
at = 0
for b0, b1 in blocks:
    shuffled[at:at + b1 - b0] = f.buf[b0:b1]
    at += b1 - b0
torch.cuda.synchronize()
for rep in range(2):
    out = kit.sort_bed_device(shuffled.data_ptr(), f.nbytes, on_device=True)
    print("rows", n, "in", f.nbytes, f.rows, "out", out.nbytes, out.rows, flush=True)
    b = device_text_to_tensor(kit, torch, out)
    a = f.buf[:f.nbytes]
    m = min(a.numel(), b.numel())
    first = None
    step = 1 << 28
    for o in range(0, m, step):
        d = (a[o:o + step][:min(step, m - o)] != b[o:o + step][:min(step, m - o)])
        if bool(d.any()):
            first = o + int(torch.nonzero(d)[0])
            break
    print("first difference at", first, flush=True)
    if first is not None:
        lo = max(0, first - 300)
        print("WANT:", bytes(a[lo:first + 300].cpu().numpy()))
        print("GOT :", bytes(b[lo:first + 300].cpu().numpy()))
    out.free()
    del b
