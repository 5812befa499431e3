"""GPU, opt-in (``SAE_TEST_EXPERIMENTAL=1``): kernels written after the round's GPU budget was spent and therefore not yet
validated on hardware.  Each runs the ordinary conv parity suite in a subprocess with the kernel's opt-in switch set (the
library reads its switches once per process)."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SAE_TEST_EXPERIMENTAL") != "1", reason="experimental kernels are opt-in")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_conv_tc6_persistent_per_tap_kernel():
    """conv_tc6_kernel (SAE_TC6=1): stride-2 fprop and small maps through the persistent, double-buffered variant of
    conv_tc2 — the stride-2 and small-map entries of CONV_CASES plus the full-network goldens"""
    env = dict(os.environ, SAE_TC6="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q",
                          "-x", "--timeout", "200", "-k", "conv or networks or layers or train_steps"],
                         env=env, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]


def test_merged_stride2_dgrad_kernel():
    """conv_tc5m_kernel (SAE_DGRAD_MERGED=1): the four parity classes of a stride-2 data gradient / transposed conv in one
    launch — stride-2 entries of CONV_CASES (dgrad direction), conv_transpose2d, layer / network goldens, training steps"""
    env = dict(os.environ, SAE_DGRAD_MERGED="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q",
                          "-x", "--timeout", "200", "-k", "conv or networks or layers or train_steps or shadow"],
                         env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
