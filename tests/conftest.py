import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_count():
    try:
        from montecarlo_simulation_b200 import api

        return api.device_count()
    except Exception:  # library not built: the gpu tests would fail loudly at import, which is what we want there
        return -1


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a CPU-only host skips the gpu-marked tests instead of erroring in their fixtures
    (the product still has no CPU fallback: these tests cannot run without a device)."""
    if _gpu_count() != 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (libmc2d has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref).  Present in the build container and, as a prebuilt
    .so shipped with the snapshot, on the GPU box; skipped if it was never built."""
    import oracle_lib

    if not oracle_lib.have_ref():
        pytest.skip("oracle/_ref/libmc2d_ref.so not built")
    return oracle_lib.Ref()


@pytest.fixture(scope="session")
def golden():
    import json

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return {
        "sanity": json.load(open(os.path.join(here, "sanity_rows.json"))),
        "ref": json.load(open(os.path.join(here, "ref_vectors.json"))),
    }
