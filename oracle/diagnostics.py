"""NumPy restatement of the reference diagnostics (autocorr / iat / ess / rhat).

TEST INFRASTRUCTURE (see oracle/__init__.py).  One 1-D chain per call, like
the reference; ``*_batch`` helpers loop over a leading series axis.
"""
from __future__ import annotations

import numpy as np


def autocorr(chain):
    """Biased sample autocorrelation at lags 0..N-1 via a zero-padded FFT.

    Follows autocorr.py:6-33: FFT length ``2**ceil(log2(2N-1))`` (:26),
    population variance ddof=0 (:27), demean (:28), power spectrum through
    ``abs()**2`` (:29-30), inverse transform scaled by ``1/var/N`` (:32).
    """
    x = np.asarray(chain)
    n = len(x)
    if n < 2:
        raise ValueError(f"autocorr requires len(chain) >= 2, but len(chain)={n}")
    nfft = 2 ** np.ceil(np.log2(2 * n - 1)).astype("int")
    v = np.var(x)
    spec = np.fft.fft(x - np.mean(x), nfft)
    power = np.abs(spec) ** 2
    return (np.fft.ifft(power).real / v / n)[:n]


def end_pos_pairs(acor):
    """First even index whose pair sum is negative, else the largest even
    index reached (iat.py:37-43)."""
    n_pairs = len(acor) // 2
    for j in range(n_pairs):
        if acor[2 * j] + acor[2 * j + 1] < 0:
            return 2 * j
    return 2 * n_pairs


def iat_ipse(chain):
    """Initial positive sequence estimator: ``2 sum_{k<n} rho_k - 1``
    (iat.py:46-92)."""
    if len(chain) < 4:
        raise ValueError(f"iat requires len(chain) >= 4, but len(chain)={len(chain)}")
    ac = autocorr(chain)
    n = end_pos_pairs(ac)
    return 2 * ac[:n].sum() - 1


def iat_imse(chain):
    """Initial monotone sequence estimator (iat.py:95-135): pair sums clipped
    by their running minimum; the first pair always counts (:127-128)."""
    if len(chain) < 4:
        raise ValueError(f"iat requires len(chain) >= 4, but len(chain)={len(chain)}")
    ac = autocorr(chain)
    n = end_pos_pairs(ac)
    low = ac[0] + ac[1]
    total = low
    for j in range(1, n // 2):
        low = min(low, ac[2 * j] + ac[2 * j + 1])
        total += low
    return 2 * total - 1


iat = iat_imse  # iat.py:138-156


def ess_ipse(chain):
    """ess.py:5-21."""
    if len(chain) < 4:
        raise ValueError("ess_ipse requires len(chain) >= 4")
    return len(chain) / iat_ipse(chain)


def ess_imse(chain):
    """ess.py:24-49."""
    if len(chain) < 4:
        raise ValueError("ess_imse requires len(chain) >= 4")
    return len(chain) / iat_imse(chain)


ess = ess_imse  # ess.py:52-69


def rhat(chains):
    """Potential scale reduction (rhat.py:111-171): ddof=1 for the within-chain
    variances and for the variance of chain means; ragged chains enter through
    the mean length (:163-170)."""
    if len(chains) < 2:
        raise ValueError("rhat requires len(chains) >= 2")
    if any(len(c) < 2 for c in chains):
        raise ValueError("rhat requires len(chain) >= 2 for every chain")
    nbar = np.mean([len(c) for c in chains])
    mu = [np.mean(c) for c in chains]
    s2 = [np.var(c, ddof=1) for c in chains]
    return np.sqrt((nbar - 1) / nbar + np.var(mu, ddof=1) / np.mean(s2))


def split_chains(chains):
    """rhat.py:9-24 (first half one longer for odd sizes)."""
    out = []
    for c in chains:
        c = np.asarray(c)
        h = (len(c) + 1) // 2
        out.extend([c[:h], c[h:]])
    return out


def split_rhat(chains):
    """rhat.py:174-202."""
    return rhat(split_chains(chains))


def rank_chains(chains):
    """rhat.py:27-59: ascending ranks from 1 over the concatenation of all chains
    (``argsort().argsort() + 1``), handed back chain by chain.  Ties: the
    reference's default (unstable) sort leaves their order implementation
    defined; a STABLE sort is used here (ties ranked in flattened order) -- the
    rule the device sort follows too.  No reference test exercises ties."""
    if len(chains) == 0:
        return chains
    flat = np.concatenate([np.asarray(c, dtype=np.float64) for c in chains])
    ranks = (flat.argsort(kind="stable").argsort(kind="stable") + 1).astype(np.float64)
    out, i = [], 0
    for c in chains:
        out.append(ranks[i:i + len(c)])
        i += len(c)
    return out


def rank_normalize_chains(chains):
    """rhat.py:62-108: ``norm.ppf((rank - 0.325) / (S - 0.25))`` with S the total
    number of draws (the code's constants; its docstring says 3/8)."""
    from scipy.special import ndtri   # == scipy.stats.norm.ppf for loc 0, scale 1
    S = sum(len(c) for c in chains)
    return [ndtri((r - 0.325) / (S - 0.25)) for r in rank_chains(chains)]


def rank_normalized_rhat(chains):
    """rhat.py:205-236."""
    return split_rhat(rank_normalize_chains(chains))


# ---- batch helpers ---------------------------------------------------------
def autocorr_batch(x):
    """x [S, N] -> [S, N]."""
    return np.stack([autocorr(r) for r in x])


def iat_ess_batch(x, estimator="imse"):
    f = iat_imse if estimator == "imse" else iat_ipse
    t = np.array([f(r) for r in x])
    return t, x.shape[1] / t


def rhat_batch(x):
    """x [chains, draws, params] -> [params]."""
    return np.array([rhat(list(x[:, :, p])) for p in range(x.shape[2])])


def sample_ar1(phi, n, rng):
    """AR(1) generator of test_iat.py:11-15 (x_t = phi x_{t-1} + e_t)."""
    z = rng.normal(size=n)
    for t in range(1, n):
        z[t] += phi * z[t - 1]
    return z
