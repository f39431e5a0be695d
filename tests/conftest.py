import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=[pytest.param("cuda", marks=pytest.mark.gpu), pytest.param("emul")])
def engine_backend(request):
    """The two ways the parity suites run the engine: "cuda" = the product library on a B200
    (gpu-marked); "emul" = the same kernel sources compiled for the CPU SIMT emulator under
    tests/native/emul (test infrastructure; runs in the CPU suite)."""
    if request.param == "emul":
        if request.node.get_closest_marker("gpu"):
            pytest.skip("CUDA-only test")
        import emul_lib
        with emul_lib.use():
            yield "emul"
    else:
        yield "cuda"
