"""GPU: handle lifetimes through the C ABI.  Maps keep their context alive (include/hpp_b200.h, hpp_ctx_destroy), so a
garbage-collected binding may finalise a context before its maps -- which CPython does for objects caught in a reference
cycle at interpreter shutdown (weak references are cleared before the finalisers run, PEP 442)."""
import subprocess
import sys
import os

import numpy as np
import pytest

from tests.common import block_map, random_poses

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_context_destroyed_before_its_map():
    import hybrid_path_planning_b200 as hpp
    occ = block_map(64, 48, 0.05, 2, seed=1)
    ctx = hpp.Context(0)
    m = hpp.Map(ctx, occ, 0.2)
    poses = random_poses(100, 64, 48, margin=1.0, seed=2)
    want = m.collide_circles(poses)
    L, h_ctx = ctx.L, ctx.h
    ctx.h = None                       # the Python object no longer owns the handle
    assert L.hpp_ctx_destroy(h_ctx) == 0
    assert np.array_equal(m.collide_circles(poses), want)   # the map still works: it holds the context
    m.close()                          # the last reference: stream and scratch go now


def test_cycle_collected_at_interpreter_exit():
    code = """
import sys
sys.path.insert(0, %r)
import numpy as np
import hybrid_path_planning_b200 as hpp
from tests.common import block_map
class Holder:
    pass
a, b = Holder(), Holder()
a.other, b.other = b, a                # a reference cycle: collected by the GC at shutdown, finalisers in any order
a.ctx = hpp.Context(0)
b.maps = [hpp.Map(a.ctx, block_map(64, 48, 0.05, 2, seed=k), 0.2) for k in range(3)]
a.ctx_again = a.ctx
print("ok", flush=True)
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stderr[-2000:])
