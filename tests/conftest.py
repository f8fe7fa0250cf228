import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: CPU test that builds the full 185 M-parameter model")


@pytest.fixture
def fake_kernels(monkeypatch):
    """Host-logic tests: swap the kernel wrappers for their torch-CPU semantic emulation."""
    from audioldm_with_lora_b200 import ops
    from tests import fake_ops
    fake_ops.install(monkeypatch, ops)
    return fake_ops
