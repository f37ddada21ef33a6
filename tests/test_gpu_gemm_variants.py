"""The tile-scheduling variants of the plain-TF32 GEMM that the automatic dispatch only uses for some shapes -- the
persistent kernel (two TMEM accumulator stages, 1 or 2 CTAs per SM) and the persistent 2 x 2 multicast-cluster kernel --
forced onto EVERY shape of the integer-exactness matrix of tests/test_gpu_ops.py.  The switches (AIR_TC_PERSIST,
AIR_TC_CLUSTER4) are read once per process, so every variant runs in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
from tests.test_gpu_ops import _tf32_case
shapes = [(128, 128, 32), (256, 256, 256), (4096, 1024, 256), (200, 100, 50), (130, 70, 36), (2500, 1024, 512),
          (4096, 320, 256), (12288, 512, 784), (12288, 784, 512), (1000, 520, 200)]
n = 0
for M, N, Kd in shapes:
    for tA, tB in ((False, False), (False, True), (True, False), (True, True)):
        got, want = _tf32_case(M, N, Kd, tA, tB, seed=M + N + Kd, use_cinit=(N %% 2 == 0), use_bias=True, mode="tf32")
        assert np.array_equal(got, want.astype(np.float32)), (M, N, Kd, tA, tB, float(np.abs(got - want).max()))
        n += 1
print("exact on", n, "cases")
""" % ROOT


@pytest.mark.parametrize("env", [{"AIR_TC_PERSIST": "1"}, {"AIR_TC_PERSIST": "2"}, {"AIR_TC_CLUSTER4": "1"}],
                         ids=["persistent-1-per-sm", "persistent-2-per-sm", "cluster-2x2-multicast"])
def test_forced_gemm_variant_is_exact_on_integers(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=e, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "exact on 40 cases" in r.stdout
