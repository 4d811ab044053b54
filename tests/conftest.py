import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """libgort.so and the CLI mirrors are build artefacts (git-ignored): build them in-tree when a fresh checkout runs the tests."""
    import subprocess
    pkg = os.path.join(ROOT, "concurrent-raytracer-go_b200")
    want = [os.path.join(pkg, "lib", "libgort.so"), os.path.join(pkg, "bin", "raytracer"), os.path.join(pkg, "bin", "benchmark")]
    if not all(os.path.exists(w) for w in want):
        subprocess.check_call(["make", "-C", pkg, "-j4"])


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


@pytest.fixture(scope="session")
def gort(built):
    import importlib
    return importlib.import_module("concurrent-raytracer-go_b200")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O
