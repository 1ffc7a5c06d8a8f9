"""The C ABI from a pure-C consumer: tests/c_abi/abi_check.c is compiled with gcc -std=c99 -pedantic against
include/starky_bn254_b200.h and linked with the in-tree library.  Without a device it checks the record layouts, the config,
sbn_air_info and the loud no-fallback failure; with a device (`-m gpu`) it also proves a ModularStark trace end to end."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(sbn, tmp_path):
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(sbn.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-o", exe, "-L", libdir, "-lstarkybn254_b200", "-Wl,-rpath," + libdir])
    return exe


def test_c_consumer_without_device(sbn, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (covered by the gpu-marked test)")
    out = subprocess.run([_build(sbn, tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_consumer_proves_on_device(sbn, tmp_path):
    out = subprocess.run([_build(sbn, tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "byte proof" in out.stdout
