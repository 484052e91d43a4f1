"""Run the UNMODIFIED reference under recording RNG proxies.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Normals (ziggurat) and uniforms share one PCG64 bit stream with
data-dependent interleaving, so "pre-drawn stream
injection" is implemented by *recording*: the reference sampler's ``_rng`` is
wrapped, the sampler is run, and the recorded per-draw standard normals and
uniforms are what the oracle restatement and the CUDA kernels are fed.

The reference is looked up at run time (``/root/reference`` in the authoring
container, else ``baseline/_ref``); it is never vendored into this repo.
"""
from __future__ import annotations

import importlib
import os
import sys
from unittest import mock

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ["/root/reference", os.path.join(_HERE, "..", "baseline", "_ref")]


def reference_available() -> bool:
    return _find() is not None


def _find():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "bayes_kit", "hmc.py")):
            return os.path.abspath(c)
    return None


def load_reference():
    """Import the reference ``bayes_kit`` package (unmodified)."""
    root = _find()
    if root is None:
        raise RuntimeError("reference bayes_kit not found in " + ", ".join(_CANDIDATES))
    if "bayes_kit" in sys.modules:
        return sys.modules["bayes_kit"]
    sys.dont_write_bytecode = True  # /root/reference is read-only
    sys.path.insert(0, root)
    try:
        return importlib.import_module("bayes_kit")
    finally:
        sys.path.remove(root)


class RecordingGenerator:
    """Proxy for ``np.random.Generator`` exposing what the samplers call
    (``normal``, ``uniform``) and logging, per call, the *standard* normal
    vector (``normal(loc, scale, size) == loc + scale * standard_normal(size)``
    bit-exactly) and the uniform."""

    def __init__(self, seed):
        self._g = np.random.default_rng(seed)
        self.normals = []
        self.uniforms = []

    def normal(self, loc=0.0, scale=1.0, size=None):
        if size is None:
            size = np.broadcast(loc, scale).shape or None
        z = self._g.standard_normal(size)
        self.normals.append(np.array(z, copy=True))
        return loc + scale * z

    def uniform(self):
        u = self._g.random()
        self.uniforms.append(u)
        return u


def record_chain(make_sampler, n_draws, seed):
    """make_sampler(rng_proxy) -> reference sampler whose ``_rng`` will be the
    proxy.  Returns dict(draws, logps, normals[list per draw], uniforms[list
    per draw], ctor_normals)."""
    proxy = RecordingGenerator(seed)
    sampler = make_sampler()
    sampler._rng = proxy
    draws, logps, zs, us, moved = [], [], [], [], []
    for _ in range(n_draws):
        proxy.normals, proxy.uniforms = [], []
        before = sampler._theta
        th, lp = sampler.sample()
        # every reference sampler REBINDS _theta on accept (hmc.py:61,
        # mala.py:62, metropolis.py:122, drghmc.py:379)
        moved.append(sampler._theta is not before)
        draws.append(np.array(th, dtype=np.float64, copy=True))
        logps.append(float(lp))
        zs.append(list(proxy.normals))
        us.append(list(proxy.uniforms))
    return dict(draws=np.stack(draws), logps=np.array(logps), normals=zs,
                uniforms=us, sampler=sampler, accepts=np.array(moved))


class LegacyRecorder:
    """Patches the process-global legacy ``np.random`` entry points used by
    smc.py (:73 ``choice``, :81 ``normal``, :85 ``uniform``) with wrappers that
    draw from the same legacy stream and log what was consumed."""

    def __init__(self):
        self.normals, self.uniforms, self.res_uniforms, self.indices = [], [], [], []

    def __enter__(self):
        rs = np.random

        def normal(loc=0.0, scale=1.0, size=None):
            shape = np.shape(loc) if size is None else size
            z = rs.standard_normal(shape)
            self.normals.append(np.array(z, copy=True))
            return loc + scale * z

        def uniform():
            u = rs.random_sample()
            self.uniforms.append(u)
            return u

        def choice(a, size=None, replace=True, p=None):
            assert replace and p is not None and a == size
            st = rs.get_state()
            want = _orig_choice(a, size=size, replace=True, p=p)
            rs.set_state(st)
            u = rs.random_sample(size)
            cdf = np.cumsum(p)
            cdf /= cdf[-1]
            idx = cdf.searchsorted(u, side="right")
            assert np.array_equal(idx, want), "legacy choice restatement drifted"
            self.res_uniforms.append(u)
            self.indices.append(np.asarray(idx, dtype=np.int64))
            return idx

        _orig_choice = rs.choice
        self._patches = [mock.patch.object(rs, "normal", normal),
                         mock.patch.object(rs, "uniform", uniform),
                         mock.patch.object(rs, "choice", choice)]
        for p_ in self._patches:
            p_.start()
        return self

    def __exit__(self, *exc):
        for p_ in self._patches:
            p_.stop()
        return False
