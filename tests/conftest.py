import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def scale_close(a, b, tol=1e-5, msg=""):
    """'1e-5 relative' as SURVEY.md defines it: max|a-b| <= tol * max(|b|, tiny)."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f"{msg}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if a.numel() == 0:
        return
    both_nan = torch.isnan(a) & torch.isnan(b)
    a, b = torch.where(both_nan, torch.zeros_like(a), a), torch.where(both_nan, torch.zeros_like(b), b)
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    if _TOL_REPORT is not None and tol > 1e-5:   # MNR_TOL_REPORT=<file>: log how much of every tolerance above 1e-5 is used
        with open(_TOL_REPORT, "a") as f:
            f.write(f"{os.environ.get('PYTEST_CURRENT_TEST', '?').split(' ')[0]}\t{msg}\ttol={tol:g}\tmeasured={err / scale:.3e}\n")
    assert err <= tol * scale, f"{msg}: max|diff|={err:.3e} > {tol:g} * scale({scale:.3e})"


_TOL_REPORT = os.environ.get("MNR_TOL_REPORT")
