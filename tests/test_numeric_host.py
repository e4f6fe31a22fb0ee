"""CPU test: the device parser's decimal->double code (bedops_b200/csrc/strtod_exact.cuh) is compiled as plain C++
and compared bit-for-bit with glibc strtod on ~1M random and adversarial literals (tests/native/strtod_check.cpp)."""
import os
import subprocess

from conftest import ROOT


def test_device_strtod_source_matches_glibc(tmp_path):
    exe = str(tmp_path / "strtod_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "strtod_check.cpp"), "-lm"],
                   check=True)
    p = subprocess.run([exe, "1000000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 mismatches" in p.stdout


def test_device_fixed_formatter_source_matches_glibc_printf(tmp_path):
    """bedops_b200/csrc/fixed_exact.cuh ("%.<prec>f" on the device) vs printf on ~1M doubles, prec 0..18."""
    exe = str(tmp_path / "fixed_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "fixed_check.cpp"), "-lm"],
                   check=True)
    p = subprocess.run([exe, "1000000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 mismatches" in p.stdout


def test_device_sci_formatter_source_matches_glibc_printf(tmp_path):
    """to_sci in bedops_b200/csrc/fixed_exact.cuh ("%.<prec>e" on the device, --sci) vs printf, prec 0..17."""
    exe = str(tmp_path / "sci_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "sci_check.cpp"), "-lm"],
                   check=True)
    p = subprocess.run([exe, "1000000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 mismatches" in p.stdout
