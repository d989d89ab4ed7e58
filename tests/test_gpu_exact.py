"""The GPU exact first-passage sampler (ddm_simulate_exact) against the reference's own samples of
pyhddmjagsutils.simulratcliff (tests/golden/simulratcliff_samples.npz, generated from the unmodified reference by
tests/golden/make_golden_ratcliff.py) and against the Navarro-Fuss first-passage law.  Distribution-level parity
(SURVEY section 8a, row a14): the sampler has no time step, so no discretisation correction is involved."""
import os

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT
from oracle import wfpt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim():
    import bayesflow_nddms_b200 as pkg
    s = pkg.DDMSimulator(0, seed=77)
    yield s
    s.close()


def _golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "simulratcliff_samples.npz"))
    names = sorted({k.split("__")[0] for k in z.files})
    return {n: (z[n + "__params"], z[n + "__y"]) for n in names}


@pytest.mark.parametrize("name", sorted(_golden()))
def test_two_sample_ks_against_reference_sampler(sim, name):
    (A, T, Nu, B, Eta, V), y_ref = _golden()[name]
    y = sim.simulate_exact([A, T, Nu, B, 0.0, 0.0, Eta, V], 200_000, dataset_offset=0)[0]
    assert np.all(np.isfinite(y)) and np.all(np.abs(y) >= T)
    assert stats.ks_2samp(y, y_ref).pvalue > 0.01
    assert abs((y > 0).mean() - (y_ref > 0).mean()) < 4 * np.sqrt(0.25 / y_ref.size)


@pytest.mark.parametrize("params", [(1.5, 0.4, 3.0, 0.5, 1.0), (1.2, 0.35, -1.0, 0.4, 1.2), (2.0, 0.2, 0.0, 0.7, 0.8),
                                    (0.8, 0.3, 4.5, 0.25, 1.5)])
def test_one_sample_ks_against_wiener_first_passage_law(sim, params):
    """Eta = 0: the signed response times follow the Wiener first-passage law exactly (no barrier correction)."""
    A, T, Nu, B, V = params
    n = 1_000_000
    y = sim.simulate_exact([A, T, Nu, B, 0.0, 0.0, 0.0, V], n, dataset_offset=3)[0]
    pu = wfpt.ddm_prob_upper(Nu, A, B, V)
    assert abs((y > 0).mean() - pu) < 4.5 * np.sqrt(pu * (1 - pu) / n) + 1e-9
    for sign, p_side in ((1, pu), (-1, 1 - pu)):
        rt = np.abs(y[np.sign(y) == sign]) - T
        if rt.size < 2000:
            continue
        grid = np.quantile(rt, np.linspace(0.01, 0.99, 60))
        cdf = wfpt.ddm_cdf(grid, sign, Nu, A, B, V) / p_side
        emp = np.searchsorted(np.sort(rt), grid, side="right") / rt.size
        assert np.max(np.abs(emp - cdf)) < 1.63 / np.sqrt(rt.size) + 2e-4      # KS critical value at alpha = .01


def test_variability_ranges_batching_and_reproducibility(sim):
    from bayesflow_nddms_b200 import pyhddmjagsutils as m

    y = m.simulratcliff(N=50_000, Alpha=1.2, Tau=.4, Nu=1.0, Beta=.5, rangeTau=.2, rangeBeta=.4, Eta=.5, Varsigma=1.0,
                        simulator=sim, dataset_offset=11)
    assert y.shape == (50_000,) and y.dtype == np.float64
    assert np.all(np.abs(y) > .3)                          # Tau - rangeTau / 2 plus a positive decision time
    again = m.simulratcliff(N=50_000, Alpha=1.2, Tau=.4, Nu=1.0, Beta=.5, rangeTau=.2, rangeBeta=.4, Eta=.5, Varsigma=1.0,
                            simulator=sim, dataset_offset=11)
    assert np.array_equal(y, again)                        # (seed, dataset index) fixes the stream
    P = np.array([[1.2, .4, 1.0, .5, .2, .4, .5, 1.0], [1.5, .3, -2.0, .6, 0, 0, 0, 1.3]])
    both = m.batch_simulratcliff(P, 50_000, simulator=sim, dataset_offset=11)
    assert both.shape == (2, 50_000) and np.array_equal(both[0], y)
    assert abs((both[1] > 0).mean() - wfpt.ddm_prob_upper(-2.0, 1.5, .6, 1.3)) < 0.01
    assert sim.last_stats()["n_upper"] == int((both > 0).sum())
    # Nu beyond +-5 is clipped (pyhddmjagsutils.py:104-105); start on a boundary is absorbed at once
    a = sim.simulate_exact([1.0, .3, 9.0, .5, 0, 0, 0, 1.0], 20_000, dataset_offset=0)
    b = sim.simulate_exact([1.0, .3, 5.0, .5, 0, 0, 0, 1.0], 20_000, dataset_offset=0)
    assert np.array_equal(a, b)
    edge = sim.simulate_exact([1.0, .3, 1.0, 1.0, 0, 0, 0, 1.0], 100, dataset_offset=0)
    assert np.all(edge == .3)
    with pytest.raises(ValueError):
        sim.simulate_exact([1.0, .3, 1.0, .9, 0, .4, 0, 1.0], 10)       # start point range leaves [0, 1]
    with pytest.raises(ValueError):
        sim.simulate_exact([0.0, .3, 1.0, .5, 0, 0, 0, 1.0], 10)        # Alpha must be positive
