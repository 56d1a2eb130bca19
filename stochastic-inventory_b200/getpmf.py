"""Host-side mirror of the reference's pmf-table builder.

Follows /root/reference/src/sdp/inventory/GetPmf.java:82-134 (`getpmf()`): truncated, renormalised
demand tables `pmf[t][j] = (d_j, p_j)`.  The reference builds them with SSJ 3.3.0 probability
classes (pom.xml:25-29), which are not available here; scipy.stats stands in for `prob`, `cdf` and
`inverseF`, so tables agree with SSJ to rounding (~1e-15 relative), not bit for bit.  In a Java
deployment SSJ keeps producing the table and only the table crosses the C-ABI, so solver parity is
defined at the table boundary.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import stats


class Distribution:
    discrete_int = False

    def cdf(self, x):
        return float(self._d.cdf(x))

    def inverseF(self, u):
        return float(self._d.ppf(u))


class PoissonDist(Distribution):
    """umontreal.ssj.probdist.PoissonDist(lambda)."""
    discrete_int = True

    def __init__(self, mean):
        self.mean = float(mean)
        self._d = stats.poisson(self.mean)

    def prob(self, j):
        return float(self._d.pmf(j))


class NormalDist(Distribution):
    """umontreal.ssj.probdist.NormalDist(mu, sigma)."""

    def __init__(self, mu, sigma):
        self._d = stats.norm(loc=mu, scale=sigma)


class GammaDist(Distribution):
    """umontreal.ssj.probdist.GammaDist(alpha, lambda): shape alpha, rate lambda."""

    def __init__(self, alpha, lam):
        self._d = stats.gamma(a=alpha, scale=1.0 / lam)


class UniformIntDist(Distribution):
    """umontreal.ssj.probdist.UniformIntDist(i, j)."""
    discrete_int = True

    def __init__(self, i, j):
        self.i, self.j = int(i), int(j)
        self._d = stats.randint(self.i, self.j + 1)

    def prob(self, x):
        return 1.0 / (self.j - self.i + 1) if self.i <= x <= self.j else 0.0


class DiscreteDistribution(Distribution):
    """umontreal.ssj.probdist.DiscreteDistribution(values, probs, n): goes through the
    cdf-difference branch of GetPmf (GetPmf.java:120-129), giving tables with zero-probability
    support points."""

    def __init__(self, values, probs, n=None):
        order = np.argsort(values)
        self.values = np.asarray(values, dtype=float)[order]
        self.probs = np.asarray(probs, dtype=float)[order]

    def cdf(self, x):
        return float(self.probs[self.values <= x].sum())

    def inverseF(self, u):
        c = np.cumsum(self.probs)
        k = int(np.searchsorted(c, u, side="left"))
        return float(self.values[min(k, len(self.values) - 1)])


class GetPmf:
    """`new GetPmf(distributions, truncationQuantile, stepSize).getpmf()` (GetPmf.java:35-41,82-134)."""

    def __init__(self, distributions, truncationQuantile, stepSize):
        self.distributions = list(distributions)
        self.truncationQuantile = float(truncationQuantile)
        self.stepSize = float(stepSize)

    def getpmf(self):
        dists, q, step = self.distributions, self.truncationQuantile, self.stepSize
        T = len(dists)
        first = dists[0]
        if isinstance(first, UniformIntDist):  # GetPmf.java:97-111 (uses distributions[0] for every period)
            rows = []
            for _ in range(T):
                rows.append(np.array([[j, first.prob(j)] for j in range(first.i, first.j + 1)], dtype=float))
            return rows
        rows = []
        for i in range(T):
            lb = float(int(dists[i].inverseF(1 - q)))
            if first.discrete_int:
                lb = 0.0
            ub = float(int(dists[i].inverseF(q)))
            n = int((ub - lb + 1) / step)
            row = np.zeros((n, 2))
            for j in range(n):
                row[j, 0] = lb + j * step
                if first.discrete_int:
                    psum = dists[i].cdf(ub) - dists[i].cdf(lb - 1)
                    row[j, 1] = dists[i].prob(j) / psum
                else:
                    psum = dists[i].cdf(ub + 0.5 * step) - dists[i].cdf(lb - 0.5 * step)
                    row[j, 1] = (dists[i].cdf(row[j, 0] + 0.5 * step) - dists[i].cdf(row[j, 0] - 0.5 * step)) / psum
            rows.append(row)
        return rows


def poisson_pmf(means, truncationQuantile=0.9999, stepSize=1.0):
    """Shorthand used by the drivers: Poisson demand per period, truncated at the quantile."""
    return GetPmf([PoissonDist(m) for m in means], truncationQuantile, stepSize).getpmf()


def clsp_inline_pmf(means_or_distributions, truncationQuantile=0.99999, stepSize=1.0):
    """The pmf CLSP.main builds inline (src/capacitated/CLSP.java:218-247): supports are
    inverseF(1-q)..inverseF(q) without the (int) cast or the LB=0 override of GetPmf.  Two branches, chosen by
    `distributions[0] instanceof DiscreteDistribution` (:236):
      * not a DiscreteDistribution -- which includes PoissonDist, CLSP.main's own choice: SSJ's PoissonDist extends
        DiscreteDistributionInt, a different class -- cdf differences normalised by
        cdf(UB + step/2) - cdf(LB - step/2) (:241-245);
      * a generic DiscreteDistribution: `prob(demand) / (2q - 1)` (:238-239), where SSJ's DiscreteDistribution.prob(i)
        is the probability of the i-th SUPPORT POINT, so the demand value is used as an index (my reading of SSJ
        3.3.0, which is not in the image; the quirk is kept, and the normaliser 2q - 1 is not the mass of the table).
    Accepts Poisson means or distribution objects."""
    dists = [d if isinstance(d, Distribution) else PoissonDist(d) for d in means_or_distributions]
    q = float(truncationQuantile)
    discrete_branch = isinstance(dists[0], DiscreteDistribution)
    rows = []
    for d in dists:
        lb, ub = d.inverseF(1 - q), d.inverseF(q)
        n = int((ub - lb + 1) / stepSize)
        row = np.zeros((n, 2))
        for j in range(n):
            row[j, 0] = lb + j * stepSize
            demand = int(row[j, 0])
            if discrete_branch:
                probabilitySum = 2 * q - 1
                p_i = float(d.probs[demand]) if 0 <= demand < len(d.probs) else 0.0  # prob(i): i-th support point
                row[j, 1] = p_i / probabilitySum
            else:
                probabilitySum = d.cdf(ub + 0.5 * stepSize) - d.cdf(lb - 0.5 * stepSize)
                row[j, 1] = (d.cdf(row[j, 0] + 0.5 * stepSize) - d.cdf(row[j, 0] - 0.5 * stepSize)) / probabilitySum
        rows.append(row)
    return rows


class GetPmfMulti:
    """`new GetPmfMulti(distributions, truncationQuantile, stepSize).getPmf(t)` for two products
    (src/sdp/cash/multiItem/GetPmfMulti.java:35-175): rows (demand1, demand2, prob), product 1 outermost.
    Implemented branches: DiscreteDistribution (exact products, :158-171) and PoissonDist (:101-128,
    normalised by (2q-1)^2 as the reference does)."""

    def __init__(self, distributions, truncationQuantile, stepSize):
        self.distributionGeneral = distributions  # [2][T]
        self.truncationQuantile = float(truncationQuantile)
        self.stepSize = float(stepSize)

    def getPmf(self, t):
        d1, d2 = self.distributionGeneral[0][t], self.distributionGeneral[1][t]
        q, step = self.truncationQuantile, self.stepSize
        rows = []
        if isinstance(d1, DiscreteDistribution):
            for i in range(len(d1.values)):
                for j in range(len(d2.values)):
                    rows.append((d1.values[i], d2.values[j], d1.probs[i] * d2.probs[j]))
            return np.array(rows, dtype=float)
        if isinstance(d1, PoissonDist):
            lb = [float(int(d.inverseF(1 - q))) for d in (d1, d2)]
            ub = [float(int(d.inverseF(q))) for d in (d1, d2)]
            n1, n2 = int((ub[0] - lb[0] + 1) / step), int((ub[1] - lb[1] + 1) / step)
            psum = (2 * q - 1) * (2 * q - 1)
            for i in range(n1):
                for j in range(n2):
                    a, b = lb[0] + i * step, lb[1] + j * step
                    rows.append((a, b, d1.prob(int(a)) * d2.prob(int(b)) / psum))
            return np.array(rows, dtype=float)
        raise NotImplementedError("GetPmfMulti: only the discrete and Poisson branches are mirrored")

    def tables(self):
        return [self.getPmf(t) for t in range(len(self.distributionGeneral[0]))]
