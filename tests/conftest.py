"""Pytest configuration: registers the `gpu` marker and shared fixtures."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache: dict = {}

    def get(name: str) -> dict:
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def rel_err(a, b) -> float:
    """max |a-b| / max |b| (scale-relative error, robust to entries that are ~0)."""
    a, b = np.asarray(a), np.asarray(b)
    den = float(np.max(np.abs(b))) or 1.0
    return float(np.max(np.abs(a - b))) / den


def assert_elementwise(a, ref, rtol: float = 1e-9, atol_scale: float = 1e-12) -> None:
    """Every entry: |a - ref| <= rtol |ref| + atol_scale max|ref| (rel_err alone leaves small entries unconstrained)."""
    a, ref = np.asarray(a), np.asarray(ref)
    bound = rtol * np.abs(ref) + atol_scale * (float(np.max(np.abs(ref))) or 1.0)
    excess = np.abs(a - ref) - bound
    worst = int(np.argmax(excess))
    assert excess.flat[worst] <= 0, (
        f"entry {worst}: got {a.flat[worst]!r}, reference {ref.flat[worst]!r}, |Δ| = {abs(a.flat[worst] - ref.flat[worst]):.3e} "
        f"> bound {bound.flat[worst]:.3e}")
