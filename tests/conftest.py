import os
import sys

import pytest

os.environ.setdefault("OMP_NUM_THREADS", "2")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(autouse=True)
def _clean_parallel_state():
    yield
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel import schedules

    if ps.model_parallel_is_initialized():
        ps.destroy_model_parallel()
    schedules.set_schedule_override(None)
