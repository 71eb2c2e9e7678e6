import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative tolerance' of BASELINE.json: scale = the reference tensor)."""
    import torch
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def grads_close(got, want, tol, what="", gmax=None):
    """Compare dicts of gradients.  Each entry is scaled by max(|want_k|, 1e-3 * max_k |want_k|):
    gradients that are mathematically zero (e.g. the last CPB bias: a constant added to every logit
    of a softmax row) are rounding noise in the reference too and only need to be small."""
    import torch
    if gmax is None:
        gmax = max(float(v.abs().max()) for v in want.values())
    for k, w in want.items():
        g = got[k]
        assert g is not None, f"{what}{k}: no gradient"
        scale = max(float(w.abs().max()), 1e-3 * gmax, 1e-30)
        err = float((g.detach().double().cpu() - w.double().cpu()).abs().max()) / scale
        assert err < tol, f"{what}{k}: rel err {err:.3e} >= {tol}"
