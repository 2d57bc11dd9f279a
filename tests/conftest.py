import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    # the CPU-side libraries (asset packer, oracle) are cheap to build; the CUDA library is
    # built by __graft_entry__.build() and only *loaded* here
    from unitygaussiansplatting_b200 import build
    build.build_asset()
    import subprocess
    subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def g():
    import unitygaussiansplatting_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def O():
    from oracle import gs_oracle_py
    return gs_oracle_py


def default_camera(g, width=256, height=256, fov=39.09651, pos=(0.0, 0.5, -6.0), forward=(0.0, 0.0, 1.0)):
    return g.Camera(position=np.array(pos, np.float64), rotation=g.look_rotation(forward), fieldOfView=fov, pixelWidth=width,
                    pixelHeight=height)


@pytest.fixture(scope="session")
def ctx(g):
    if not has_cuda():
        pytest.skip("no CUDA device")
    return g.GaussianSplatContext(0)
