import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_built():
    import oracle_lib

    if not os.path.exists(os.path.join(oracle_lib.ORACLE_DIR, "libmsm_oracle.so")):
        oracle_lib.build_oracle()
    return oracle_lib


@pytest.fixture(scope="session")
def product_lib():
    """The CUDA library must exist (built by __graft_entry__.build()); loading needs no GPU."""
    import msm_blst_b200 as M

    if not os.path.exists(M.LIB_PATH):
        M.build_library()
    M.lib()
    return M
