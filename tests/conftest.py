import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

_guard = {"ctx": None, "checks": 0}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(autouse=True)
def _device_guard_zones(request):
    """guard mode (KX_GUARD=1, see tests/test_gpu_guard.py): after every GPU test no device buffer of the library may have
    been overrun — kx_debug_check_guards looks at every guarded buffer of the process"""
    yield
    if not os.environ.get("KX_GUARD") or request.node.get_closest_marker("gpu") is None:
        return
    import knoxdb_b200 as kb
    if _guard["ctx"] is None:
        _guard["ctx"] = kb.Context(0)
    bad = _guard["ctx"].check_guards()
    _guard["checks"] += 1
    assert bad == 0, f"{bad} device buffer(s) were written past their end during {request.node.nodeid}"


def pytest_terminal_summary(terminalreporter):
    if os.environ.get("KX_GUARD"):
        terminalreporter.write_line(f"{_guard['checks']} guard zones checked after GPU tests: all intact")
