import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu under gpurun")


@pytest.fixture
def emu_backend():
    """Inject the CPU emulation of the op contracts (tests/emu_ops.py) for host-logic tests; restore afterwards."""
    from emu_ops import EmuOps
    from polyp_image_generator_b200 import ops as ops_mod
    prev = ops_mod._backend
    ops_mod.set_backend(EmuOps())
    yield ops_mod.get()
    ops_mod.set_backend(prev)


@pytest.fixture(scope="session")
def built_lib():
    from polyp_image_generator_b200 import build as _build
    return _build.build()
