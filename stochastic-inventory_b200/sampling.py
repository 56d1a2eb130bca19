"""Host-side mirror of the reference's sample generator, src/sdp/sampling/Sampling.java:26-335.

The reference draws through SSJ 3.3.0's `umontreal.ssj.rng.MRG32k3a` (pom.xml:25-29; Sampling.java:28
`static RandomStream stream = new MRG32k3a()`), which is absent from /root/reference.  `MRG32k3a` below
restates the published generator (L'Ecuyer, "Good parameters and implementations for combined multiple
recursive random number generators", Oper. Res. 47(1), 1999; stream / substream structure of L'Ecuyer,
Simard, Chen, Kelton, "An object-oriented random-number package with many long streams and substreams",
Oper. Res. 50(6), 2002): package seed (12345 x 6), streams 2^127 apart, substreams 2^76 apart.  It is
pinned by the known first outputs of the first stream given in the second paper's example program
(0.1270111501, 0.3185275653, 0.3091860155 -- tests/test_sampling.py).

Two of the reference's generators (`generateLHSamples`, Sampling.java:86-103 and :157-177) take their
in-stratum offsets from `Math.random()`, so the reference's own simulated means change from run to run;
the mirror draws those offsets from the stream as well (as `generateLHSamples2`, :131-147, does), which
is the reproducible behaviour SURVEY.md section 8(f) asks for.  `inverseF` comes from getpmf.py (scipy).
"""
from __future__ import annotations

import numpy as np

_M1 = 4294967087
_M2 = 4294944443
_A12, _A13N = 1403580, 810728
_A21, _A23N = 527612, 1370589
_NORM = 2.328306549295727688e-10  # 1 / (m1 + 1)

# one-step transition matrices of the two components (state vectors are column vectors (s0, s1, s2))
_A1 = ((0, 1, 0), (0, 0, 1), (-_A13N % _M1, _A12, 0))
_A2 = ((0, 1, 0), (0, 0, 1), (-_A23N % _M2, 0, _A21))


def _matmul(a, b, m):
    return tuple(tuple(sum(a[i][k] * b[k][j] for k in range(3)) % m for j in range(3)) for i in range(3))


def _matpow2(a, e, m):
    """a^(2^e) mod m."""
    for _ in range(e):
        a = _matmul(a, a, m)
    return a


def _matvec(a, v, m):
    return [sum(a[i][k] * v[k] for k in range(3)) % m for i in range(3)]


_A1P76, _A2P76 = _matpow2(_A1, 76, _M1), _matpow2(_A2, 76, _M2)
_A1P127, _A2P127 = _matpow2(_A1, 127, _M1), _matpow2(_A2, 127, _M2)


class MRG32k3a:
    """`new MRG32k3a()`: each new object is the next stream of the package (2^127 steps further on)."""
    _next_seed = [12345] * 6

    @classmethod
    def setPackageSeed(cls, seed):
        seed = [int(x) for x in seed]
        if len(seed) != 6 or not all(0 <= s < _M1 for s in seed[:3]) or not all(0 <= s < _M2 for s in seed[3:]) \
                or not any(seed[:3]) or not any(seed[3:]):
            raise ValueError("MRG32k3a seed: six integers, first three < m1, last three < m2, neither triple all zero")
        cls._next_seed = seed

    def __init__(self):
        cls = type(self)
        self.Ig = list(cls._next_seed)  # start of the stream
        self.Bg = list(self.Ig)         # start of the current substream
        self.Cg = list(self.Ig)         # current state
        cls._next_seed = _matvec(_A1P127, self.Ig[:3], _M1) + _matvec(_A2P127, self.Ig[3:], _M2)

    def resetStartStream(self):
        self.Bg = list(self.Ig)
        self.Cg = list(self.Ig)

    def resetStartSubstream(self):
        self.Cg = list(self.Bg)

    def resetNextSubstream(self):
        self.Bg = _matvec(_A1P76, self.Bg[:3], _M1) + _matvec(_A2P76, self.Bg[3:], _M2)
        self.Cg = list(self.Bg)

    def nextDouble(self):
        s = self.Cg
        p1 = (_A12 * s[1] - _A13N * s[0]) % _M1
        s[0], s[1], s[2] = s[1], s[2], p1
        p2 = (_A21 * s[5] - _A23N * s[3]) % _M2
        s[3], s[4], s[5] = s[4], s[5], p2
        return (p1 - p2) * _NORM if p1 > p2 else (p1 - p2 + _M1) * _NORM

    def nextInt(self, i, j):
        """RandomStream.nextInt(i, j): uniform over {i, ..., j}."""
        return i + int(self.nextDouble() * (j - i + 1.0))


class Sampling:
    """Sampling.java:26-335.  The reference shares one static stream between all instances; so does this."""
    stream = None

    def __init__(self):
        if Sampling.stream is None:
            Sampling.stream = MRG32k3a()

    @staticmethod
    def resetStartStream():
        Sampling.stream.resetStartStream()

    @staticmethod
    def resetNextSubstream():
        Sampling.stream.resetNextSubstream()

    def generateRanSamples(self, distributions, sampleNum):
        """Sampling.java:52-63: plain Monte Carlo, period-major draw order."""
        T = len(distributions)
        out = np.empty((sampleNum, T))
        for i in range(T):
            for j in range(sampleNum):
                out[j, i] = distributions[i].inverseF(0.0 + (1.0 - 0.0) * self.stream.nextDouble())
        return out

    def getNextSample(self, distributions):
        """Sampling.java:72-80."""
        return np.array([d.inverseF(self.stream.nextDouble()) for d in distributions])

    def generateLHSamples2(self, distributions, sampleNum):
        """Sampling.java:131-147 (and :86-103 with the stream in place of Math.random()): next substream, one
        stratum [j/n, (j+1)/n) per sample and period, inverse cdf, column shuffle."""
        T = len(distributions)
        out = np.empty((sampleNum, T))
        self.resetNextSubstream()
        for i in range(T):
            for j in range(sampleNum):
                randomNum = 0.0 + (1.0 / sampleNum - 0.0) * self.stream.nextDouble()
                lowBound = j / sampleNum
                out[j, i] = distributions[i].inverseF(lowBound + randomNum)
        return self.shuffle(out)

    generateLHSamples = generateLHSamples2

    def shuffle(self, samples):
        """Sampling.java:325-334: for every column, swap each row with a uniformly drawn row."""
        n = samples.shape[0]
        for i in range(samples.shape[1]):
            for j in range(n):
                mark = self.stream.nextInt(0, n - 1)
                samples[j, i], samples[mark, i] = samples[mark, i], samples[j, i]
        return samples
