"""pytest configuration: `gpu` marker, repo root on sys.path, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; compiled on first use)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engine():
    """A live Engine on cuda:0.  No fallback: a missing library or GPU is an error, not a skip."""
    from khoice_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
