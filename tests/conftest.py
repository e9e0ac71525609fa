import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
