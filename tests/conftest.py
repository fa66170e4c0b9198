import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle_lib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libtrajref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from trajectory_generator_ros2_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(params=["fast", "exact"])
def mode_engine(request, engine):
    """The engine in each planning mode (tgx_set_plan_mode): fast exact-v jumps (default) and step-by-step replay."""
    engine.set_plan_mode(request.param == "exact")
    yield engine
    engine.set_plan_mode(False)
