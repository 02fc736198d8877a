import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(autouse=True)
def _fresh_fp16_range_state():
    """The fp16 range guard of the no-grad path is process-wide and sticky (ops.INFER_DTYPE falls back to bf16 once an
    activation exceeded +-65504): every test starts from the configured format and a clear flag."""
    import sys
    ops = sys.modules.get("probabilistic_domain_adaptation_b200.ops")
    if ops is None:
        yield
        return
    configured = getattr(ops, "_CONFIGURED_INFER_DTYPE", None)
    if configured is None:
        configured = ops._CONFIGURED_INFER_DTYPE = ops.INFER_DTYPE
    ops.INFER_DTYPE = configured
    for st in ops._RANGE.values():
        st["flag"].zero_()
        st["pending"] = False
    yield
