"""Stretcher (reference: bayes_kit/ensemble.py).

Upstream the class body is a docstring only -- the whole Goodman & Weare
stretch-move implementation is commented out (ensemble.py:16-66) -- so there is
nothing executable to match; the name is exported for surface compatibility.
"""


class Stretcher:
    """Affine-invariant ensemble sampler placeholder (not executable upstream either)."""
