import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the CPU oracle (torch/oneDNN) is fastest at <= 16 threads; on the 128-core GPU host the default thread count
    # oversubscribes and is two orders of magnitude slower (profiles/r01_cpu_threads_sweep.txt)
    try:
        import torch
        torch.set_num_threads(min(16, os.cpu_count() or 1))
    except Exception:
        pass


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _no_shim_leak():
    """oracle/tf_shim.py installs a fake `tensorflow` in sys.modules; never let it outlive the test that asked for it."""
    yield
    shim = sys.modules.get("oracle.tf_shim")
    if shim is not None:
        shim.uninstall()
