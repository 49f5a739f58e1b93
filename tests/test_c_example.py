"""The C ABI is usable from plain C: examples/hal_roundtrip.c compiles with gcc -std=c99 against include/r0b200.h and
links libr0b200.so (CPU); on a GPU box it runs and reports a clean round trip."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = "/tmp/r0b200_hal_roundtrip"


def _compile():
    libdir = os.path.join(ROOT, "risc0_b200", "lib")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "hal_roundtrip.c"), "-L", libdir, "-lr0b200", "-Wl,-rpath," + libdir, "-o", EXE]
    subprocess.check_call(cmd)


def test_c_example_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_c_example_runs():
    _compile()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("ok:")
