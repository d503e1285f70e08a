import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def O():
    return entry.load_oracle()


@pytest.fixture(scope="session")
def lib(pkg):
    return pkg.load_library()


@pytest.fixture(scope="session")
def oracle_c():
    import ctypes

    path = os.path.join(ROOT, "oracle", "_ref", "liboracle_c.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/liboracle_c.so not built (run __graft_entry__.build())")
    return ctypes.CDLL(path)


@pytest.fixture(scope="session")
def ref_lib():
    """The reference's own CUDA path, rebuilt unmodified (oracle/ref_harness.cu)."""
    import ctypes

    path = os.path.join(ROOT, "oracle", "_ref", "libsfm_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsfm_ref.so not built (needs /root/reference at build time)")
    L = ctypes.CDLL(path)
    L.ref_create.restype = ctypes.c_void_p
    for name in ("ref_estimateE", "ref_computePosecandidates", "ref_choosePose", "ref_linear_triangulation",
                 "ref_estimateE_injected", "ref_host_det"):
        getattr(L, name).restype = ctypes.c_float
    return L


@pytest.fixture(scope="session")
def scene_small(O):
    K, Kinv = O.reference_K()
    sc = O.synthetic_pair(3000, seed=4321)
    sc["K"], sc["Kinv"] = K, Kinv
    sc["x"] = O.normalise_points(sc["px"], Kinv)
    return sc
