"""Overrun detection without compute-sanitizer (closed on this pool): with KX_GUARD=1 the library allocates every device
scratch / result buffer exactly as large as needed, followed by 256 bytes of 0xFA, and tests/conftest.py checks every zone
after each test.  Together with the 32 guard bytes behind every host output the binding allocates (knoxdb_b200/lib.py,
mirroring internal/cmp/tests/gen.go:13-44) a kernel or copy that writes past the end of a buffer fails the run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECTION = ("simple8b or string_in_sets or bitset_golden or bitset_ops or multi_predicate or predicate_trees or row_masks or alprd_blocks or float32_raw or scan_select or "
             "time_bucketed or string_blocks or alp_float or run_end or in_sets or plan or pipeline or batching or sharded_on_one_rank or cmp_random or "
             "gather_bytes or scan_host_takes or zone_maps")


def test_selected_gpu_tests_leave_every_guard_zone_intact():
    if os.environ.get("KX_GUARD"):
        pytest.skip("already inside the guarded run")
    env = dict(os.environ, KX_GUARD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-q", "-x", "-k", SELECTION, "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "guard zones checked" in r.stdout, tail
