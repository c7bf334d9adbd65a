import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Builds the native libraries in-tree (no-op when they are current)."""
    import importlib

    build = importlib.import_module("raytracing-1w_b200.build")
    build.build_all()
    return True


@pytest.fixture(scope="session")
def rt(built):
    import importlib

    return importlib.import_module("raytracing-1w_b200")


@pytest.fixture(scope="session")
def oracle(built):
    import oracle_binding

    oracle_binding.load()
    return oracle_binding


@pytest.fixture(scope="session")
def gpu_ctx(rt):
    ctx = rt.Context(0)
    yield ctx
    ctx.close()
