import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a hung kernel must fail the test, not hang the GPU box until the outer limit
    for item in items:
        if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(300, method="thread"))


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The .so is git-ignored: on a fresh checkout build it (nvcc cross-compiles without a GPU) so that the
    C-ABI surface tests do not depend on the order in which the driver runs build() and the tests."""
    from bayesflow_nddms_b200 import _build

    if not os.path.exists(_build.LIB_PATH):
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: the tests that need the library will say so
            print(f"could not build libddm_b200.so: {e}")
    yield


@pytest.fixture(scope="session")
def golden():
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu

    cpu.build()
    return cpu


@pytest.fixture(scope="session")
def sim():
    """One simulator for the whole GPU session.  Fails (does not skip) if the CUDA library
    cannot be used: a GPU test that silently fell back would prove nothing."""
    import bayesflow_nddms_b200 as pkg

    s = pkg.DDMSimulator(device=0, seed=2023)
    yield s
    s.close()


VARIANT_MODEL = {"basic": 0, "alpha": 1, "alpha_dc": 2, "alpha_scale": 3, "alpha_scale2": 4, "alpha_fine": 1, "stahl": 5, "eta": 6}


def variant_kwargs(variant):
    if variant == "alpha_fine":
        return dict(dt=0.001, max_steps=4000)
    return dict(dt=0.01, max_steps=400)
