import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One hpp_ctx on cuda:0 for the whole GPU session.  No skip-on-failure: a missing library or GPU must fail
    loudly under -m gpu (there is no CPU fallback to fall back to)."""
    import hybrid_path_planning_b200 as hpp
    c = hpp.Context(0)
    yield c
    c.close()
