import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("aa"))


@pytest.fixture(scope="session")
def product_lib():
    """The C-ABI library must exist: build it (nvcc) when missing — never fall back to anything else."""
    import alignasm_b200 as aa
    if not os.path.exists(aa.lib_path()):
        import __graft_entry__ as g
        g.build()
    return aa.load_library()
