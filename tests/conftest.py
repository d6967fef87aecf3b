"""pytest configuration: the `gpu` marker, import paths, and loaders for the golden vectors."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).resolve().parent / "golden"
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _native_library_built():
    """The in-tree librtb.so is a build product (git-ignored): build it if a fresh checkout does not have it yet."""
    import subprocess
    lib = ROOT / "ray_trace_pb_b200" / "_lib" / "librtb.so"
    if not lib.exists():
        subprocess.run(["make", "-C", str(ROOT / "ray_trace_pb_b200" / "csrc"), "-j4"], check=True,
                       stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def rt():
    import ray_trace_pb_b200.raytrace as rt
    return rt


@pytest.fixture(scope="session")
def rtm():
    import ray_trace_pb_b200.materials as rtm
    return rtm


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle
    oracle.build()
    return oracle


def load_golden(name):
    return np.load(GOLDEN / f"{name}.npz", allow_pickle=False)


@pytest.fixture(scope="session")
def checksums():
    return json.loads((GOLDEN / "checksums.json").read_text())


@pytest.fixture(scope="session")
def host_api():
    return json.loads((GOLDEN / "host_api.json").read_text())
