"""pytest configuration: `gpu` marker, import paths, one-time build of the C-ABI library."""

from __future__ import annotations

import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (or reuse) conch_b200/libconch_b200.so.  nvcc cross-compiles sm_100a without a GPU."""
    from conch_b200.build import build

    return build()


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
