"""pytest configuration: `gpu` marker, import paths, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lidar-rt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
BG = np.array([0.0, 0.0, 1.0], np.float32)     # (intensity, hit, drop) background, train.py:104-106


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle
    return Oracle(False)


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle
    return Oracle(True)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def scene_args(sc):
    """(means, scales, rots, opac, shs) from a synthetic.Scene or a dict/npz."""
    if hasattr(sc, "means"):
        return sc.means, sc.scales, sc.rots, sc.opac, sc.shs
    return sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"]


def assert_close(a, b, atol, rtol, what=""):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    err = np.abs(a - b); tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} outside tol; max err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)} (ref {b.flat[err.argmax()]:.6g})"


def grad_close(a, b, rel, what=""):
    """Gradient comparison relative to the tensor's max magnitude (atomics / cancellation)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-12)
    err = np.abs(a - b).max() / scale
    assert err <= rel, f"{what}: max err / max|ref| = {err:.3e} > {rel}"
