"""Out-of-bounds writes: every kernel family on small meshes with guard words behind every device block.

compute-sanitizer is not available on the pool; the library's own guard mode (NPB_DEBUG_GUARDS=1, csrc/capi.cu) is the
substitute.  The switch is read once per process, hence the subprocess."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_no_kernel_writes_past_its_buffers():
    env = dict(os.environ, NPB_DEBUG_GUARDS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "guard_suite.py"), "quick"], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "guard suite done" in r.stdout
    assert " damaged 0" in r.stdout and "guards 0 " not in r.stdout
