"""The persistent multi-step STEP launch (all interior leapfrog steps of a draw in one kernel, clusters ordered by
global release/acquire counters) must produce exactly the draws of one launch per step: the arithmetic is identical,
only the launch structure differs.  BK_TC_FUSE is read once per process, so each arm runs in its own interpreter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(shape, fuse):
    env = dict(os.environ, BK_TC_FUSE=fuse)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuse_check.py"), *map(str, shape)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-1]


@pytest.mark.parametrize("shape", [(2048, 1000, 6), (777, 300, 5), (9472, 1000, 10)])
def test_fused_steps_equal_step_launches(shape):
    fused, stepwise = _run(shape, "1"), _run(shape, "0")
    assert fused == stepwise
    digest, lp_digest, accept = fused.split()
    assert 0.5 < float(accept) <= 1.0


def test_two_streams_do_not_interleave_fused_launches():
    """Two samplers driven from two streams: their persistent multi-step launches each fill the GPU, so the library
    chains them behind one another (an event recorded after every fused launch); the draws equal the sequential ones."""
    import numpy as np
    import torch

    import bayes_kit_b200 as bk
    from oracle.models import DensePrecGauss

    model = bk.DensePrecGauss(DensePrecGauss.c2_precision(512, 0), dtype=torch.float32)

    def run(concurrent):
        a = bk.HMCDiag(model, 0.1, 6, chains=16384, seed=1)
        b = bk.HMCDiag(model, 0.1, 6, chains=16384, seed=2)
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        outs = []
        for _ in range(3):
            for smp, st in ((a, sa), (b, sb)):
                with torch.cuda.stream(st if concurrent else torch.cuda.current_stream()):
                    outs.append(smp.sample()[0])
        torch.cuda.synchronize()
        return [o.cpu().numpy() for o in outs]

    for x, y in zip(run(True), run(False)):
        np.testing.assert_array_equal(x, y)
