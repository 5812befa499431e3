"""GPU: alternative kernel selections, each run against the ordinary conv parity suite in a subprocess with the selecting
switch set (the library reads its switches once per process).  The DEFAULT selection is covered by the main suite; this file
keeps the alternatives honest: ``SAE_DGRAD_MERGED=0`` (stride-2 data gradient as one launch per parity class),
``SAE_DISABLE_TCGEN05=1`` (every convolution on the generic mma.sync kernel), ``SAE_FUSED_BLOCKS=0`` (per-operator autograd
nodes in the discriminators' ResBlocks).  Opt-in (``SAE_TEST_EXPERIMENTAL=1``) because each case repeats a few minutes of tests."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SAE_TEST_EXPERIMENTAL") != "1", reason="alternative kernel selections are opt-in")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"SAE_DGRAD_MERGED": "0"}, {"SAE_DGRAD_MERGED": "1"}, {"SAE_WGRAD_PAIR": "0"}, {"SAE_DISABLE_TCGEN05": "1"}, {"SAE_FUSED_BLOCKS": "0"}],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
def test_alternative_kernel_selection(env):
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q",
                          "-x", "--timeout", "300", "-k", "conv or networks or layers or train_steps or shadow"],
                         env=dict(os.environ, **env), capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
