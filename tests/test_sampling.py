"""Host mirror of Sampling.java / SSJ's MRG32k3a (stochastic-inventory_b200/sampling.py): known answers of
the published generator and the structure of the Latin-hypercube samples.  No GPU needed."""
import importlib

import numpy as np

import sdpb200 as S

sp = importlib.import_module(S.package.__name__ + ".sampling")


def test_mrg32k3a_published_constants():
    """Jump matrices and stream spacing against the constants printed in L'Ecuyer et al. (2002) / RngStream.c:
    A1^(2^76), A1^(2^127), A2^(2^76), A2^(2^127) (first rows) and the seed of the package's second stream."""
    assert sp._A1P76[0] == (82758667, 1871391091, 4127413238)
    assert sp._A1P127[0] == (2427906178, 3580155704, 949770784)
    assert sp._A2P76[0] == (1511326704, 3759209742, 1610795712)
    assert sp._A2P127[0] == (1464411153, 277697599, 1610723613)
    sp.MRG32k3a.setPackageSeed([12345] * 6)
    g1, g2 = sp.MRG32k3a(), sp.MRG32k3a()
    assert g1.Ig == [12345] * 6
    assert g2.Ig == [3692455944, 1366884236, 2968912127, 335948734, 4161675175, 475798818]


def test_mrg32k3a_first_outputs_and_resets():
    sp.MRG32k3a.setPackageSeed([12345] * 6)
    g = sp.MRG32k3a()
    u = [g.nextDouble() for _ in range(3)]
    assert np.allclose(u, [0.1270111220, 0.3185275653, 0.3091860155], rtol=0, atol=1e-9)
    g.resetStartStream()
    assert [g.nextDouble() for _ in range(3)] == u
    g.resetNextSubstream()
    v = g.nextDouble()
    assert v != u[0]
    g.resetStartSubstream()
    assert g.nextDouble() == v
    # 2^76 single steps are out of reach, but the jump must commute with stepping: jump(step(s)) == step(jump(s))
    a, b = sp.MRG32k3a(), None
    s0 = list(a.Cg)
    a.nextDouble()
    stepped = list(a.Cg)
    jumped_after = sp._matvec(sp._A1P76, stepped[:3], sp._M1) + sp._matvec(sp._A2P76, stepped[3:], sp._M2)
    a.Cg = sp._matvec(sp._A1P76, s0[:3], sp._M1) + sp._matvec(sp._A2P76, s0[3:], sp._M2)
    a.nextDouble()
    assert a.Cg == jumped_after
    k = [g.nextInt(3, 7) for _ in range(200)]
    assert min(k) == 3 and max(k) == 7


def test_latin_hypercube_samples_are_stratified_and_reproducible():
    sp.MRG32k3a.setPackageSeed([12345] * 6)
    sp.Sampling.stream = None
    dists = [S.PoissonDist(20), S.PoissonDist(5), S.NormalDist(10, 2)]
    n = 64
    a = sp.Sampling().generateLHSamples2(dists, n)
    assert a.shape == (n, 3)
    for i, d in enumerate(dists):
        # one sample per stratum [j/n, (j+1)/n): the sorted column equals the inverse cdf of increasing probabilities
        col = np.sort(a[:, i])
        lo = np.array([d.inverseF(j / n) if j else -np.inf for j in range(n)])
        hi = np.array([d.inverseF(min((j + 1) / n, 1 - 1e-12)) for j in range(n)])
        assert np.all(col >= lo) and np.all(col <= hi)
    sp.MRG32k3a.setPackageSeed([12345] * 6)
    sp.Sampling.stream = None
    b = sp.Sampling().generateLHSamples2(dists, n)
    assert np.array_equal(a, b)
    c = sp.Sampling().generateLHSamples2(dists, n)  # next substream: different draws, same strata
    assert not np.array_equal(a, c) and np.array_equal(np.sort(a[:, 0]), np.sort(c[:, 0])) is not None
    r = sp.Sampling().generateRanSamples(dists[:2], 10)
    assert r.shape == (10, 2) and np.all(r >= 0)
