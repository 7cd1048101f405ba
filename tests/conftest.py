import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# Order of the GPU suite under `-x`: building blocks first, then memory safety, then short parity, then
# the long jobs -- a failure in a long test must not hide the cheap ones (round-1 lesson).
_ORDER = [("test_gpu_tcgen05", 0), ("test_kernels_write_only_inside_their_buffers", 1),
          ("test_full_config1_1000_steps", 8), ("test_config3_", 8), ("test_full_pair_list", 9), ("test_gpu_dropin", 7)]


def _rank(item):
    for key, r in _ORDER:
        if key in item.nodeid:
            return r
    return 5


def pytest_collection_modifyitems(session, config, items):
    items.sort(key=_rank)   # stable: file order is kept inside a rank


@pytest.fixture(scope="session")
def built_lib():
    """libvlg_b200.so, built in-tree with nvcc (cross-compiles without a GPU)."""
    import vlg_b200
    return vlg_b200.build.build()


@pytest.fixture(scope="session")
def selftest_lib():
    """libvlg_b200_selftest.so: the tcgen05 building-block kernels, tests only."""
    import vlg_b200
    from vlg_b200 import _lib
    vlg_b200.build.build_selftest()
    return _lib.load_selftest()
