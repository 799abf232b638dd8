import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "cpp-11-ray-trace-march-framework_b200"


def pkg(sub=None):
    return importlib.import_module(PKG + ("." + sub if sub else ""))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: long-running")


def _have_gpu():
    try:
        lib = pkg("capi").load_library()
        return lib.cuda_trace_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; when they are selected there must be a GPU -- they
    # then FAIL (not skip) if the CUDA library cannot be loaded, so a silent fallback is impossible
    pass


@pytest.fixture(scope="session")
def pyoracle():
    from oracle import pyoracle as po
    po.build(ref=True, port=True)
    return po


@pytest.fixture(scope="session")
def port(pyoracle):
    return pyoracle.Port.get()


@pytest.fixture(scope="session")
def ref(pyoracle):
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs /root/reference)")
    return pyoracle.Ref.get()


class SceneData:
    """Post-transform mesh arrays + camera of one preset, built through a MeshApi backend."""

    def __init__(self, name, vtx, tri, fov, cam16):
        self.name, self.vtx, self.tri, self.fov, self.cam16 = name, vtx, tri, float(fov), cam16


_scene_cache = {}


@pytest.fixture(scope="session")
def scene_data(pyoracle):
    """scene_data(name) -> SceneData.  Built with the reference's own Mesh/Matrix44f when
    oracle/_ref exists, else with the host library (whose parity with the reference is pinned by
    tests/test_host_vs_ref.py in the build container)."""
    scenes = pkg("scenes")

    def get(name):
        if name not in _scene_cache:
            if pyoracle.have_ref():
                api = pyoracle.Ref.get().api
            else:
                api = pkg("hostapi").host_api()
            m, fov, cam = scenes.build(api, name)
            vtx, tri = m.arrays()
            _scene_cache[name] = SceneData(name, vtx, tri, fov, cam)
        return _scene_cache[name]

    return get


@pytest.fixture(scope="session")
def cuda_trace():
    capi = pkg("capi")
    ct = capi.CudaTrace(1)
    yield ct
    ct.close()
