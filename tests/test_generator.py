"""The production normal generator, quantified.

The stepping kernels draw normals from 21-bit Box-Muller fields: 2^21 radii, |z| <= 5.5226, where the reference
uses full-range fp64 normals.  The CPU half computes the map's exact law (oracle/normal_map.py) and bounds its
distance from N(0, 1); the GPU half checks >= 1e10 normals of the device map against that exact law."""
import math

import numpy as np
import pytest


def test_discrete_map_law_is_close_to_normal():
    from oracle import normal_map as nm

    r = nm.radii()
    assert abs(nm.Z_CAP - 5.5225425) < 1e-6 and abs(r.max() - nm.Z_CAP) < 1e-12
    assert 3.2e-8 < nm.lost_tail_mass() < 3.5e-8                  # mass of N(0,1) beyond the cap, per normal
    edges = np.linspace(0.0, 5.6, 113)
    pm, pn, tail_map, tail_norm = nm.bin_probabilities(edges, r)
    assert abs(tail_map) < 1e-12 and abs(pm.sum() - 1.0) < 1e-12
    rel = np.abs(pm - pn) / pn
    lo = edges[:-1]
    # bins of 0.05 sigma: indistinguishable from N(0,1) to 3.5 sigma, within 1.5e-4 to 4.4, 1.1 % to 4.8, 11 % to 5.2; beyond
    # that the eight radii out there make the law lumpy (total mass 1e-7)
    assert rel[lo < 3.0].max() < 1e-6 and rel[lo < 4.0].max() < 5e-5 and rel[lo < 4.4].max() < 3e-4
    assert rel[lo < 4.8].max() < 0.02 and rel[lo < 5.2].max() < 0.12
    # Kolmogorov distance and binned total variation (0.02 grid): what bounds any event's probability per draw
    grid = np.arange(0.0, 6.0, 0.02)
    cm, cn = nm.abs_cdf(grid, r), nm.normal_abs_cdf(grid)
    ks = float(np.max(np.abs(cm - cn)))
    tv = 0.5 * float(np.abs(np.diff(cm) - np.diff(cn)).sum() + abs((1 - cm[-1]) - (1 - cn[-1])))
    assert ks < 1e-7 and tv < 5e-7
    # consequence for a first-passage event of an n-step trial: |P_map(A) - P_normal(A)| <= n * tv
    assert 4000 * tv < 2e-3 and 400 * tv < 2e-4
    # second and fourth moments of the map (exact in the radius): E z^2 = E r^2 / 2, E z^4 = 3/8 E r^4
    assert abs(np.mean(r ** 2) / 2 - 1.0) < 1e-5 and abs(3.0 / 8.0 * np.mean(r ** 4) - 3.0) < 2e-4


@pytest.mark.gpu
def test_device_normals_match_the_map_law_at_1e10(sim):
    from oracle import normal_map as nm

    r = nm.radii()
    edges = np.linspace(0.0, 5.6, 113)
    pm, pn, _, _ = nm.bin_probabilities(edges, r)

    # (1) chi-square of 2.4e9 normals against the map's exact law (at this n the MUFU approximation errors, ~1e-6
    # absolute in z, still move less mass across a bin edge than the bin's own sampling noise)
    g = sim.normals_histogram(2_400_000_000, 112, 5.6, 256, seed=11)
    n = g["n"]
    assert int(g["abs"].sum()) + g["beyond"] == n and g["beyond"] == 0
    keep = pm * n >= 50
    chi2 = float((((g["abs"][keep] - pm[keep] * n) ** 2) / (pm[keep] * n)).sum())
    dof = int(keep.sum()) - 1
    z = (chi2 - dof) / math.sqrt(2 * dof)
    print(f"|z| histogram, n = {n:.3g}: chi2 = {chi2:.1f} on {dof} dof ({z:+.2f} sigma)")
    assert z < 4.5
    ang = g["angle"].astype(np.float64)
    e = ang.sum() / ang.size
    chi2a = float(((ang - e) ** 2 / e).sum())
    za = (chi2a - (ang.size - 1)) / math.sqrt(2 * (ang.size - 1))
    print(f"pair angles in {ang.size} sectors: chi2 = {chi2a:.1f} ({za:+.2f} sigma)")
    assert int(ang.sum()) == n // 2 and za < 4.5
    m1, m2, m3, m4 = g["moments"]
    assert abs(m1) < 5 / math.sqrt(n) and abs(m2 - 1) < 5 * math.sqrt(2 / n) + 2e-5
    assert abs(m3) < 5 * math.sqrt(15 / n) and abs(m4 - 3) < 5 * math.sqrt(96 / n) + 3e-4

    # (2) the tails at 1.2e10 normals: every bin from 4 sigma out agrees with the map's law (Poisson 5 sigma + 1 % for the
    # MUFU rounding of the few radii out there), nothing beyond the cap, and the count beyond 5 sigma is what the
    # 2^-21-spaced radii give -- about 2 % above the normal's, not 8 000 missing values' worth of distortion elsewhere
    t = sim.normals_histogram(12_000_000_000, 112, 5.6, 256, seed=12)
    n = t["n"]
    assert t["beyond"] == 0
    lo = edges[:-1] >= 4.0
    exp = pm[lo] * n
    got = t["abs"][lo].astype(np.float64)
    assert np.all(np.abs(got - exp) <= 5 * np.sqrt(exp) + 0.01 * exp + 3), (got, exp)
    beyond5 = float(t["abs"][edges[:-1] >= 5.0].sum())
    exp5_map, exp5_norm = float(pm[edges[:-1] >= 5.0].sum()) * n, math.erfc(5 / math.sqrt(2)) * n
    print(f"n = {n:.3g}: |z| >= 5: {beyond5:.0f} (map law {exp5_map:.0f}, normal {exp5_norm:.0f}); lost beyond the cap: {nm.lost_tail_mass() * n:.0f}")
    assert abs(beyond5 - exp5_map) < 5 * math.sqrt(exp5_map) + 0.01 * exp5_map
