"""The C++ host mirror (host/GaussianSplatRenderer.hpp) compiles against the C ABI; on the GPU box it renders."""
import subprocess
from pathlib import Path

import pytest

from conftest import has_cuda

ROOT = Path(__file__).resolve().parents[1]


def _build(tmp_path):
    exe = tmp_path / "host_mirror_test"
    lib = ROOT / "unitygaussiansplatting_b200"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", str(ROOT / "tests" / "host_mirror_test.cpp"), "-o", str(exe), "-L" + str(lib),
                    "-lgsplat_b200", "-Wl,-rpath," + str(lib)], check=True)
    return exe


@pytest.mark.skipif(has_cuda(), reason="CPU-only behaviour")
def test_host_mirror_links_and_fails_loudly_without_a_device(tmp_path):
    r = subprocess.run([str(_build(tmp_path))], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_host_mirror_renders(tmp_path):
    r = subprocess.run([str(_build(tmp_path))], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
