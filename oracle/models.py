"""NumPy model classes implementing the reference model protocol.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference protocol is ``bayes_kit/typing.py:15-42``:
``dims()``, ``log_density(theta)``, ``log_density_gradient(theta) -> (lp, grad)``
and for SMC ``log_prior(theta)``, ``log_likelihood(theta)``.  The reference's
own test models are 1-D (``test/models/std_normal.py:5-13``); BASELINE.json's
model families do not exist upstream, so these classes *define* them for both
sides: the unmodified reference samplers are driven by these objects to make
the golden fixtures, and the CUDA plugins implement the same densities.
"""
from __future__ import annotations

import numpy as np


class IsoGauss:
    """D-dim isotropic Gaussian, ``log p = -0.5 * theta.theta / sigma^2``.

    D-dim generalisation of ``test/models/std_normal.py:8-13`` (sigma=1).
    """

    def __init__(self, dims: int, sigma: float = 1.0):
        self._d = int(dims)
        self.sigma = float(sigma)
        self._prec = 1.0 / (self.sigma * self.sigma)

    def dims(self) -> int:
        return self._d

    def log_density(self, theta):
        return -0.5 * (self._prec * np.dot(theta, theta))

    def log_density_gradient(self, theta):
        return self.log_density(theta), -(self._prec * theta)


class DiagGauss:
    """``log p = -0.5 * sum(prec * (theta - mu)^2)``."""

    def __init__(self, mu, prec):
        self.mu = np.asarray(mu, dtype=np.float64)
        self.prec = np.asarray(prec, dtype=np.float64)

    def dims(self) -> int:
        return self.mu.shape[0]

    def log_density(self, theta):
        r = theta - self.mu
        return -0.5 * np.dot(r, self.prec * r)

    def log_density_gradient(self, theta):
        r = theta - self.mu
        g = -(self.prec * r)
        return -0.5 * np.dot(r, self.prec * r), g


class DensePrecGauss:
    """``log p = -0.5 (theta-mu)^T P (theta-mu)``, ``grad = -P (theta-mu)``.

    BASELINE config c2: ``P = A A^T / D + I``.
    """

    def __init__(self, P, mu=None):
        self.P = np.ascontiguousarray(P, dtype=np.float64)
        self.mu = None if mu is None else np.asarray(mu, dtype=np.float64)

    @staticmethod
    def c2_precision(D: int, seed: int = 0):
        rng = np.random.default_rng(seed)
        A = rng.normal(size=(D, D))
        return A @ A.T / D + np.eye(D)

    def dims(self) -> int:
        return self.P.shape[0]

    def log_density(self, theta):
        return self.log_density_gradient(theta)[0]

    def log_density_gradient(self, theta):
        r = theta if self.mu is None else theta - self.mu
        Pr = self.P @ r
        return -0.5 * np.dot(r, Pr), -Pr


class GaussPriorLik:
    """Tempering target for SMC (BASELINE config c4).

    ``log_prior = -0.5 sum(p0 (theta-m0)^2)``,
    ``log_likelihood = -0.5 sum(pl (theta-mu)^2)``.
    c4: m0=0, p0=1, pl=4 (``ll = -2 |theta-mu|^2``), mu ~ N(0, I).
    Implements ``bayes_kit/typing.py:37-42``.
    """

    def __init__(self, m0, p0, mu, pl):
        self.m0 = np.asarray(m0, dtype=np.float64)
        self.p0 = np.asarray(p0, dtype=np.float64)
        self.mu = np.asarray(mu, dtype=np.float64)
        self.pl = np.asarray(pl, dtype=np.float64)

    def dims(self) -> int:
        return self.mu.shape[0]

    def log_prior(self, theta):
        r = theta - self.m0
        return -0.5 * np.sum(self.p0 * r * r)

    def log_likelihood(self, theta):
        r = theta - self.mu
        return -0.5 * np.sum(self.pl * r * r)

    def log_density(self, theta):
        return self.log_likelihood(theta) + self.log_prior(theta)

    def log_density_gradient(self, theta):
        g = -(self.pl * (theta - self.mu)) - self.p0 * (theta - self.m0)
        return self.log_density(theta), g

    def log_likelihood_gradient(self, theta):
        return -(self.pl * (theta - self.mu))

    def log_prior_gradient(self, theta):
        return -(self.p0 * (theta - self.m0))


class Binomial:
    """Beta-binomial on the logit scale -- the reference's own test target
    (``test/models/binomial.py:11-74``, used by test_tempered_smc.py:8-30, test_hmc.py:69,
    test_mala.py:26, test_drghmc.py:152) restated with elementary functions and an ANALYTIC
    gradient (the upstream test model differentiates numerically, binomial.py:59-65):
        p = inv_logit(theta);  log_likelihood = log C(N, x) + x log p + (N - x) log(1 - p)
        log_prior = Beta(p; alpha, beta) + log p + log(1 - p)        (logit Jacobian, binomial.py:46-50)
                  = alpha log p + beta log(1 - p) - log B(alpha, beta)
    gen_golden.py checks both densities against the upstream class (scipy) to 1e-12."""

    def __init__(self, alpha, beta, x, N):
        import math
        self.alpha, self.beta, self.x, self.N = float(alpha), float(beta), float(x), float(N)
        self.lchoose = math.lgamma(N + 1) - math.lgamma(x + 1) - math.lgamma(N - x + 1)
        self.lbeta = math.lgamma(alpha) + math.lgamma(beta) - math.lgamma(alpha + beta)

    def dims(self) -> int:
        return 1

    @staticmethod
    def _logs(th):
        e = np.exp(-np.abs(th))
        l = np.log1p(e)
        lp1 = -l if th >= 0 else th - l              # log p     = -softplus(-theta)
        l1m = -th - l if th >= 0 else -l             # log (1-p) = -softplus(theta)
        p = 1.0 / (1.0 + e) if th >= 0 else e / (1.0 + e)
        return lp1, l1m, p

    def log_likelihood(self, theta):
        lp1, l1m, _ = self._logs(float(theta[0]))
        return (self.lchoose + self.x * lp1) + (self.N - self.x) * l1m

    def log_prior(self, theta):
        lp1, l1m, _ = self._logs(float(theta[0]))
        return (self.alpha * lp1 + self.beta * l1m) - self.lbeta

    def log_density(self, theta):
        return self.log_likelihood(theta) + self.log_prior(theta)

    def log_likelihood_gradient(self, theta):
        _, _, p = self._logs(float(theta[0]))
        return np.array([self.x - self.N * p])

    def log_prior_gradient(self, theta):
        _, _, p = self._logs(float(theta[0]))
        return np.array([self.alpha - (self.alpha + self.beta) * p])

    def log_density_gradient(self, theta):
        return self.log_density(theta), self.log_likelihood_gradient(theta) + self.log_prior_gradient(theta)


class Tempered:
    """``lp_t(theta) = log_likelihood(theta) * t + log_prior(theta)`` (smc.py:47-51) as a GradModel,
    gradient ``grad_ll * t + grad_prior``: what an MCMC kernel inside the SMC targets."""

    def __init__(self, model, t):
        self.model, self.t = model, float(t)

    def dims(self) -> int:
        return self.model.dims()

    def log_density(self, theta):
        return self.model.log_likelihood(theta) * self.t + self.model.log_prior(theta)

    def log_density_gradient(self, theta):
        g = self.model.log_likelihood_gradient(theta) * self.t + self.model.log_prior_gradient(theta)
        return self.log_density(theta), g


class HierLogReg:
    """Hierarchical logistic regression (BASELINE config c3; builder-defined,
    there is no upstream definition -- SURVEY.md section 8(c) caveat 4).

    Parameters ``theta = (beta[Dx], mu, lam)`` with ``tau = exp(lam)``:
        y_n ~ Bernoulli(sigmoid(x_n . beta))
        beta_j ~ N(mu, tau^2),  mu ~ N(0, 1),  tau ~ half-N(0, 1)
    ``log p`` (constants dropped, + Jacobian ``lam`` of the log transform):
        sum_n [y_n z_n - softplus(z_n)]
        - Dx*lam - 0.5 exp(-2 lam) sum_j (beta_j - mu)^2
        - 0.5 mu^2 - 0.5 exp(2 lam) + lam
    """

    def __init__(self, X, y):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.Dx = self.X.shape[1]

    @staticmethod
    def c3_data(N: int, Dx: int, seed: int = 0):
        rng = np.random.default_rng(seed)
        X = rng.normal(size=(N, Dx)) / np.sqrt(Dx)
        beta = rng.normal(size=Dx)
        p = 1.0 / (1.0 + np.exp(-(X @ beta)))
        y = (rng.uniform(size=N) < p).astype(np.float64)
        return X, y

    def dims(self) -> int:
        return self.Dx + 2

    def log_density(self, theta):
        return self.log_density_gradient(theta)[0]

    def log_density_gradient(self, theta):
        Dx = self.Dx
        beta, mu, lam = theta[:Dx], theta[Dx], theta[Dx + 1]
        z = self.X @ beta
        # softplus(z) = max(z,0) + log1p(exp(-|z|))
        sp = np.maximum(z, 0.0) + np.log1p(np.exp(-np.abs(z)))
        ll = np.dot(self.y, z) - np.sum(sp)
        sig = 0.5 * (1.0 + np.tanh(0.5 * z))
        e2 = np.exp(-2.0 * lam)
        r = beta - mu
        ss = np.dot(r, r)
        lp = ll - Dx * lam - 0.5 * e2 * ss - 0.5 * mu * mu - 0.5 * np.exp(2.0 * lam) + lam
        g = np.empty(Dx + 2)
        g[:Dx] = self.X.T @ (self.y - sig) - e2 * r
        g[Dx] = e2 * np.sum(r) - mu
        g[Dx + 1] = -Dx + e2 * ss - np.exp(2.0 * lam) + 1.0
        return lp, g


# ---- (de)serialisation of model definitions into the golden fixtures --------
def model_spec(model) -> dict:
    """Arrays that define ``model``; stored inside each tests/golden/*.npz."""
    if isinstance(model, IsoGauss):
        return dict(model_kind="iso", model_dims=model.dims(), model_sigma=model.sigma)
    if isinstance(model, DiagGauss):
        return dict(model_kind="diag", model_mu=model.mu, model_prec=model.prec)
    if isinstance(model, DensePrecGauss):
        # P is regenerated from its seed (c2_precision) to keep fixtures small
        return dict(model_kind="dense", model_dims=model.dims(), model_seed=0,
                    model_mu=np.zeros(0) if model.mu is None else model.mu)
    if isinstance(model, GaussPriorLik):
        return dict(model_kind="gpl", model_m0=model.m0, model_p0=model.p0,
                    model_mu=model.mu, model_pl=model.pl)
    if isinstance(model, HierLogReg):
        return dict(model_kind="hlr", model_X=model.X, model_y=model.y)
    if isinstance(model, Binomial):
        return dict(model_kind="binom", model_abxn=np.array([model.alpha, model.beta, model.x, model.N]))
    raise TypeError(type(model))


def build_model(z):
    """Inverse of model_spec for a loaded npz."""
    kind = str(z["model_kind"])
    if kind == "iso":
        return IsoGauss(int(z["model_dims"]), float(z["model_sigma"]))
    if kind == "diag":
        return DiagGauss(z["model_mu"], z["model_prec"])
    if kind == "dense":
        P = DensePrecGauss.c2_precision(int(z["model_dims"]), int(z["model_seed"]))
        mu = z["model_mu"]
        return DensePrecGauss(P, None if mu.size == 0 else mu)
    if kind == "gpl":
        return GaussPriorLik(z["model_m0"], z["model_p0"], z["model_mu"], z["model_pl"])
    if kind == "hlr":
        return HierLogReg(z["model_X"], z["model_y"])
    if kind == "binom":
        a, b, x, n = z["model_abxn"]
        return Binomial(a, b, int(x), int(n))
    raise ValueError(kind)
