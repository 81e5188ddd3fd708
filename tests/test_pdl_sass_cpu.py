"""No kernel of the built library may read producer-written memory ahead of its griddepcontrol.wait.

With `const T* __restrict__` parameters nvcc emits ld.global.nc for such reads and schedules them above the wait (round 2: the
token id and position loads of embed_step_kernel; CUDA-graph replay then embedded the previous token).  common.cuh erases the
qualifier; tools/check_pdl_sass.py lists every LDG.E...CONSTANT that precedes ACQBULK in the SASS and allows only audited
constant loads (explicit __ldg of weights)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_noncoherent_load_precedes_the_dependency_wait():
    lib = os.path.join(ROOT, "stable-ts_b200", "libstablets_b200.so")
    if not os.path.exists(lib) or shutil.which("cuobjdump") is None:
        pytest.skip("library or cuobjdump not available")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_pdl_sass.py"), lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
