import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def report():
    """Appends one line per check to gpurun_out/parity_report.txt (read back after a gpurun call)."""
    outDir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(outDir, exist_ok=True)
    path = os.path.join(outDir, "parity_report.txt")
    fh = open(path, "a")

    def _write(name, **kw):
        fh.write(name + " " + " ".join("%s=%s" % (k, v) for k, v in kw.items()) + "\n")
        fh.flush()

    yield _write
    fh.close()
