import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import uvrt_testlib as T  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-size CPU checks")


@pytest.fixture(scope="session")
def uv():
    """The product package (ctypes view of libuvrt.so / libuvrt_host.so); builds it if needed."""
    mod = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
    if not (os.path.exists(os.path.join(mod.build_dir(), "libuvrt.so"))
            and os.path.exists(os.path.join(mod.build_dir(), "libuvrt_host.so"))):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def checkers():
    """Builds oracle/_build (and oracle/_ref when the reference sources are present)."""
    if not os.path.exists(os.path.join(T.ORACLE_DIR, "_build", "libuvrt_oracle.so")) or \
            (os.path.isdir("/root/reference/cl") and not T.ref_available()):
        T.build_checkers()
    return T


@pytest.fixture(scope="session")
def room(uv):
    """testroomopt.glb through the PRODUCT loader and builder: (tris, nodes, triIdx, floorHeight)."""
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    tris, nodes, tri_idx = sim.mesh_data()
    floor = sim.mesh_info()["floor"]
    sim.close()
    return tris, nodes, tri_idx, floor


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(HERE, "golden", "appendix_c.json")))


def lange_pos0(floor):
    """Lamp position of lange_route.xml:10 at lamp_hoogte 0.4."""
    f = np.float32
    return (f(-0.25500134), f(f(floor) + f(0.40000001)), f(-3.3149862))
