import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Write the measured parity errors of this session (llark_b200.testing.check_close) next to their bounds."""
    import os

    try:
        from llark_b200.testing import PARITY_LOG
    except Exception:
        return
    if not PARITY_LOG:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_dir = os.path.join(root, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "parity_measured.txt"), "w") as f:
        f.write("# measured error | bound | check (relative Frobenius unless the name says otherwise)\n")
        for name, measured, tol in PARITY_LOG:
            f.write(f"{measured:.4e} | {tol:.1e} | {name}\n")
