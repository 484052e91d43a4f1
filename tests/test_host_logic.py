"""CPU-only: the C-ABI library loads and exports every symbol include/bk.h
declares, plus the host-side logic (argument validation, sharding, accept
rules, non-plugin rejection).  No compute calls (no GPU here)."""
import math
import os
import re

import numpy as np
import pytest

import bayes_kit_b200 as bk
from bayes_kit_b200 import _lib, dist
from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "bk.h")).read()
    return sorted(set(re.findall(r"^BK_API [^\n(]*?(bk_[a-z0-9_]+)\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    decl = _declared()
    assert len(decl) >= 28
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/bk.h but not exported"
    assert sorted(_lib.EXPORTS) == decl, "ctypes signature table out of sync with include/bk.h"
    assert lib.bk_abi_version() == 1
    assert lib.bk_last_error() is not None


def test_reference_name_surface():
    # bayes_kit/__init__.py:17-30
    for name in ["DrGhmcDiag", "HMCDiag", "MALA", "Metropolis", "MetropolisHastings",
                 "TemperedLikelihoodSMC", "Stretcher", "ess", "ess_imse", "ess_ipse", "iat",
                 "iat_imse", "iat_ipse", "rhat", "autocorr"]:
        assert hasattr(bk, name)


def test_python_models_are_rejected_loudly():
    class PyModel:
        def dims(self): return 1
        def log_density(self, t): return 0.0
        def log_density_gradient(self, t): return 0.0, t
    for make in (lambda: bk.HMCDiag(PyModel(), 0.1, 3), lambda: bk.MALA(PyModel(), 0.1),
                 lambda: bk.Metropolis(PyModel(), bk.GaussianRW(1.0)),
                 lambda: bk.DrGhmcDiag(PyModel(), 1, [0.1], [1], 0.5)):
        with pytest.raises(TypeError, match="device plugin"):
            make()
    with pytest.raises(TypeError, match="proposal descriptor"):
        bk.Metropolis(None, lambda x: x)


# ---- DrGhmcDiag constructor errors: types and messages of drghmc.py:105-207,
# the regexes of test_drghmc.py:179-347 -------------------------------------------
def _dr(**kw):
    args = dict(model=None, max_proposals=3, leapfrog_step_sizes=[1.0, 0.5, 0.25],
                leapfrog_step_counts=[1, 2, 3], damping=0.2)
    args.update(kw)
    return bk.DrGhmcDiag(**args)


@pytest.mark.parametrize("kw,exc,msg", [
    (dict(max_proposals=2.0), TypeError, "max_proposals must be an int, not <class 'float'>"),
    (dict(max_proposals=0), ValueError, "max_proposals must be greater than or equal to 1, not 0"),
    (dict(leapfrog_step_sizes=1.0), TypeError, "leapfrog_step_sizes must be an instance of type sequence, but found type <class 'float'>"),
    (dict(leapfrog_step_sizes=[1.0, 0.5]), ValueError, "leapfrog_step_sizes must be a sequence of length 3, so that each proposal has its own specified leapfrog step size, but instead found length of 2"),
    (dict(leapfrog_step_sizes=[1.0, 1, 0.25]), TypeError, "each step size in leapfrog_step_sizes must be of type float, but found step size of type <class 'int'> at index 1"),
    (dict(leapfrog_step_sizes=[1.0, -0.5, 0.25]), ValueError, "each step size in leapfrog_step_sizes must be positive, but found step size of -0.5 at index 1"),
    (dict(leapfrog_step_counts=3), TypeError, "leapfrog_step_counts must be an instance of type sequence, but found type <class 'int'>"),
    (dict(leapfrog_step_counts=[1, 2]), ValueError, "leapfrog_step_counts must be a sequence of length 3, so that each proposal has its own specified number of leapfrog steps, but instead found length of 2"),
    (dict(leapfrog_step_counts=[1, 2.0, 3]), TypeError, "each step count in leapfrog_step_counts must be of type int, but found step count of type <class 'float'> at index 1"),
    (dict(leapfrog_step_counts=[1, 0, 3]), ValueError, "each step count in leapfrog_step_counts must be positive, but found step count of 0 at index 1"),
    (dict(damping=1), TypeError, "damping must be of type float, but found type <class 'int'>"),
    (dict(damping=0.0), ValueError, "damping must be within (0, 1], but found damping of 0.0"),
    (dict(damping=1.5), ValueError, "damping must be within (0, 1], but found damping of 1.5"),
])
def test_drghmc_validation(kw, exc, msg):
    with pytest.raises(exc, match=re.escape(msg)):
        _dr(**kw)


def test_accept_rules():  # test_metropolis.py:19-103 (mocked uniform)
    class U:
        def __init__(self, v): self.v = v
        def uniform(self): return self.v
    assert bk.metropolis_accept_test(math.log(0.5), math.log(0.4), U(1))
    assert not bk.metropolis_accept_test(math.log(0.4), math.log(0.8), U(0.5))
    assert bk.metropolis_accept_test(math.log(0.4), math.log(0.7999999), U(0.5))
    bal = math.log(0.5)
    assert not bk.metropolis_hastings_accept_test(math.log(0.4), math.log(0.81), bal, bal, U(0.5))
    assert bk.metropolis_hastings_accept_test(math.log(0.4), math.log(0.81), math.log(0.4), math.log(0.6), U(0.5))
    assert bk.metropolis_accept_test(-1e9, 0.0, U(0.0))  # log(0) = -inf accepts


def test_end_pos_pairs():  # test_iat.py:72-80
    from bayes_kit_b200.iat import _end_pos_pairs
    for chain, pos in [([], 0), ([1], 0), ([1, -0.5], 2), ([1, -0.5, 0.25], 2),
                       ([1, -0.5, 0.25, -0.3], 2), ([1, -0.5, 0.25, -0.1], 4),
                       ([1, -0.5, 0.25, -0.3, 0.05], 2), ([1, -0.5, 0.25, -0.1, 0.05], 4)]:
        assert _end_pos_pairs(chain) == pos


def test_shard_ranges():
    for n, w in [(10, 3), (8192, 8), (7, 8), (1_000_000, 8), (0, 2)]:
        rs = [dist.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        sizes = [hi - lo for lo, hi in rs]
        assert max(sizes) - min(sizes) <= 1 and sizes == dist.shard_sizes(n, w)


def test_no_cuda_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bk.IsoGauss(3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bk.ess(np.arange(10.0))


def test_bench_refuses_every_diagnostic_switch():
    """Every environment switch the library reads (getenv in csrc/, BK_LIB in _lib.py) is a diagnostic configuration
    bench.py must refuse to measure."""
    import glob
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    switches = {"BK_LIB"}
    for f in glob.glob(os.path.join(root, "bayes-kit_b200", "csrc", "*.cu")):
        switches |= set(re.findall(r'getenv\("(BK_[A-Z0-9_]+)"\)', open(f).read()))
    bench = open(os.path.join(root, "bench.py")).read()
    assert len(switches) >= 9
    for s in sorted(switches):
        assert f'"{s}"' in bench, f"bench.py does not refuse {s}"
