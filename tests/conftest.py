"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic, C-ABI symbol
checks, and the same kernel sources executed by the test-only SIMT emulator under tests/emu/).
`-m gpu` runs on a B200 and calls the product library through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def assets():
    from vpho_b200 import synthetic as syn
    mano = syn.make_mano_model()
    return {"mano": mano, "anchors": syn.make_anchor_assets(mano), "objects": syn.make_object_tables()}


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a GPU box (fails loudly when it is missing)."""
    import torch
    from vpho_b200 import capi
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return capi.lib()


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY: the kernel sources compiled against tests/emu/cuda_emu.h (single-threaded SIMT emulator)."""
    from tests.emu import build_emu
    from vpho_b200 import capi
    return capi.Library(build_emu.build(), strict=False)
