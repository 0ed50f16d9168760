import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _pin_exact_engine(request):
    """GPU parity tests run on the exact-fp32 FFMA engine unless they ask for the `tc` fixture
    (tests/test_gpu_tc.py), so tolerances do not depend on which engine is the package default."""
    if "gpu" in request.keywords:
        import keras_rs_b200 as K
        K.set_gemm_engine("ffma")
    yield
