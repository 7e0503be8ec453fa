"""Synthetic PSD / Balding-Nichols genotypes.

theta_n ~ Dir(0.1 * 1_K); per locus ancestral p ~ U(0.05, 0.95), beta_lk ~ Beta(p(1-F)/F,
(1-p)(1-F)/F) with F = 0.1; y ~ Binomial(2, theta_n . beta_l).  Uses numpy's legacy
RandomState so that a given seed gives the same data on every numpy version."""
import numpy as np


def psd_params(n, l, k, seed=1, f=0.1, dir_alpha=0.1):
    rs = np.random.RandomState(seed)
    theta = rs.dirichlet(np.full(k, dir_alpha), size=n)
    p = rs.uniform(0.05, 0.95, size=l)
    a = p * (1 - f) / f
    b = (1 - p) * (1 - f) / f
    beta = rs.beta(a[:, None], b[:, None], size=(l, k))
    return theta, beta


def psd_genotypes(n, l, k, seed=1, missing_rate=0.0, f=0.1):
    """-> (y[l, n] uint8 in {0,1,2,3=missing}, theta[n,k], beta[l,k]); host-side, small shapes."""
    theta, beta = psd_params(n, l, k, seed, f)
    rs = np.random.RandomState(seed + 7919)
    q = np.clip(beta @ theta.T, 0.0, 1.0)  # [l, n]
    y = rs.binomial(2, q).astype(np.uint8)
    if missing_rate > 0:
        y[rs.uniform(size=y.shape) < missing_rate] = 3
    return y, theta, beta
