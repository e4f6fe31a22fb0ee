"""Regenerates tests/golden/starch_{bz2,gz}.starch and starch_expected.bed with the reference's own `starch` / `unstarch`
(oracle/_ref/bin, built by oracle/build_ref.sh).  Run from the repository root in the build container."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bedops_b200 import synth  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
OUT = os.path.dirname(os.path.abspath(__file__))
text = synth.bed_text(3000, 11, synth.MAP_SHAPE) 
bed = os.path.join(OUT, "starch_expected.bed")
open(bed, "wb").write(text)
for flag, name in (("--bzip2", "starch_bz2.starch"), ("--gzip", "starch_gz.starch")):
    blob = subprocess.run([os.path.join(BIN, "starch"), flag, bed], capture_output=True, check=True).stdout
    open(os.path.join(OUT, name), "wb").write(blob)
    back = subprocess.run([os.path.join(BIN, "unstarch"), os.path.join(OUT, name)], capture_output=True, check=True).stdout
    assert back == text
print("ok")
