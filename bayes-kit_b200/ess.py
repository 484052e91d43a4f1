"""Effective sample size (reference: bayes_kit/ess.py): ``len(chain) / iat*(chain)``."""
from __future__ import annotations

from . import _lib as L
from .iat import _iat_ess


def ess_ipse(chain, device="cuda", draws_first=False):
    """ess.py:5-21."""
    return _iat_ess(chain, L.IAT_IPSE, device, draws_first, "ess_ipse(chain)")[1]


def ess_imse(chain, device="cuda", draws_first=False):
    """ess.py:24-49."""
    return _iat_ess(chain, L.IAT_IMSE, device, draws_first, "ess_imse(chain)")[1]


def ess(chain, device="cuda", draws_first=False):
    """ess.py:52-69 (delegates to the IMSE estimator)."""
    return _iat_ess(chain, L.IAT_IMSE, device, draws_first, "ess(chain)")[1]
