"""Drop-in for the simulation half of ``imputation_from_stahl_not_scaled.py`` (lines 49-228).

The reference standardises the Pe/c amplitudes, maps them to single-trial boundaries
``(z + 3)/3`` clipped at 0 (:82-105), draws one (drift, beta, ter, dc) per participant
(:156-177), then calls a pure-Python ``diffusion_trial`` once per CSV row (:205-213) and
stacks ``(imputed_choicert, alpha_like_Pe)`` (:228).  Here the 19 374 per-row calls are one
kernel launch (``impute_choicert``); everything else keeps the reference's arithmetic.
"""
from __future__ import annotations

import numpy as np

from . import _capi, priors
from ._model_common import configurator, device_configurator  # noqa: F401
from .priors import truncnorm_better  # noqa: F401
from .simulator import default_simulator

RNG = np.random.default_rng(2024)


def unique_inverse(ids):
    """``np.unique(ids, return_inverse=True)`` (the reference's participant index, :59-63) -- for integer ids in a
    compact range by a presence table and its prefix sum instead of a sort (0.21 -> 0.07 ms for the CSV's 19 374 rows);
    anything else goes to ``np.unique``.  Same arrays either way (tests)."""
    a = np.asarray(ids)
    if a.ndim == 1 and a.dtype.kind in "iu" and a.size:
        lo, hi = int(a.min()), int(a.max())
        if hi - lo <= 4 * a.size + 1024:
            off = a - lo if lo else a
            seen = np.zeros(hi - lo + 1, dtype=bool)
            seen[off] = True
            lut = np.cumsum(seen) - 1
            return (np.flatnonzero(seen) + lo).astype(a.dtype, copy=False), lut[off]
    return np.unique(a, return_inverse=True)


def boundaries_from_pe(all_Pe):
    """:82-105 -> (alpha_like_Pe, single_trial_alphas), both (n,) float64."""
    all_Pe = np.asarray(all_Pe, dtype=np.float64)
    # (all_Pe - mean) / std, (. + 3) / 3 twice, clip -- the reference's operations in its order, but the centred column
    # is formed once (np.std forms it again: same sums, same bits -- tests) and the two identical columns are one
    alpha_like_Pe = all_Pe - np.mean(all_Pe)
    sd = np.sqrt(np.add.reduce(alpha_like_Pe * alpha_like_Pe) / all_Pe.size) if all_Pe.size else np.float64(np.nan)
    alpha_like_Pe /= sd
    alpha_like_Pe += 3
    alpha_like_Pe /= 3
    single_trial_alphas = alpha_like_Pe.copy()
    single_trial_alphas[single_trial_alphas < 0] = 0
    return alpha_like_Pe, single_trial_alphas


def draw_participant_params(nsubs, rng=None):
    """:165-174 -> (nsubs, 4) float64 [Drift, Beta, Ter, Dc]."""
    return priors.draw_prior_batch("stahl", nsubs, RNG if rng is None else rng)


def diffusion_trial(drift, bound_trial, beta, ter, dc, dt=.01, max_steps=400., simulator=None):
    """:120-148 -> choicert.  Raises ValueError for a negative boundary, as the reference does.
    Every call consumes fresh randomness (the simulator's trial counter advances), like the reference's."""
    out = impute_choicert([0], [bound_trial], [[drift, beta, ter, dc]], dt=dt, max_steps=max_steps,
                          simulator=simulator)
    return float(out[0])


def impute_choicert(part_index, single_trial_alphas, part_params, dt=.01, max_steps=400., simulator=None,
                    seed=None, trial_offset=None, precision=32):
    """The per-row loop :205-213 as one launch.

    part_index (n,) int -- row of ``part_params`` for each trial; single_trial_alphas (n,);
    part_params (G, 4) [Drift, Beta, Ter, Dc].  Returns imputed_choicert (n,) float64.
    ``trial_offset=None``: the simulator's trial counter keys the trials and advances by n, so repeated
    imputations differ (as repeated runs of the reference's loop do); an explicit offset reproduces one."""
    sim = simulator if simulator is not None else default_simulator()
    out = sim.simulate_trialwise(part_index, single_trial_alphas, part_params, dt, int(max_steps), seed=seed,
                                 trial_offset=trial_offset, precision=precision)
    return out[:, 0].copy()


def impute_dataset(subj_idx, all_Pe, part_params=None, simulator=None, device=False, seed=None, trial_offset=None):
    """Lines 82-228 end to end on arrays: returns (input_data (n, 2), part_ids, part_index).

    ``input_data`` = column_stack((imputed_choicert, alpha_like_Pe)) (:228); with
    ``device=True`` it is a float32 torch tensor on the simulator's GPU (DLPack hand-off of
    the choicert column + the observed Pe column copied once)."""
    subj_idx = np.asarray(subj_idx)
    part_ids, part_index = unique_inverse(subj_idx)
    alpha_like_Pe, single_trial_alphas = boundaries_from_pe(all_Pe)
    if part_params is None:
        part_params = draw_participant_params(part_ids.size)
    sim = simulator if simulator is not None else default_simulator()
    if not device:
        cr = impute_choicert(part_index, single_trial_alphas, part_params, simulator=sim, seed=seed, trial_offset=trial_offset)
        return np.column_stack((cr, alpha_like_Pe)), part_ids, part_index
    import torch

    batch = sim.simulate_trialwise(part_index, single_trial_alphas, part_params, seed=seed, trial_offset=trial_offset,
                                   flags=_capi.FLAG_OUT_F32, device=True)
    t = torch.from_dlpack(batch)  # (n, 2): choicert, boundary used
    t[:, 1] = torch.as_tensor(alpha_like_Pe.astype(np.float32), device=t.device)
    return t, part_ids, part_index


def participant_batches(input_data, part_index, n_parts):
    """:235-241 -- yields the reference's per-participant ``obs_dict`` (ragged groups)."""
    for p in range(n_parts):
        these = part_index == p
        if hasattr(input_data, "device") and not isinstance(input_data, np.ndarray):
            import torch

            sub = input_data[torch.as_tensor(these, device=input_data.device)]
            yield {'sim_data': sub[None, :, :], 'sim_non_batchable_context': int(these.sum()), 'prior_draws': None}
        else:
            sub = input_data[these, ]
            yield {'sim_data': sub[np.newaxis, :, :], 'sim_non_batchable_context': int(np.sum(these)),
                   'prior_draws': None}


def load_stahl_csv(path='stahl_data/base_data.csv'):
    """:54 -- (subj_idx, pre_Pe) from the reference's CSV (needs pandas; the data is not shipped here)."""
    import pandas as pd

    df = pd.read_csv(path)
    return df['subj_idx'].to_numpy(), df['pre_Pe'].to_numpy(dtype=np.float64)


def synthetic_stahl_like(rng=None, nsubs=89, ntrials_total=19374):
    """Synthetic data with the CSV's shape (89 subjects, 13-337 trials
    each, 19 374 rows; pre_Pe mean 0, sd 5.8, range +-35) for tests and benchmarks."""
    rng = np.random.default_rng(2024) if rng is None else rng
    w = rng.uniform(13, 337, nsubs)
    counts = np.maximum(13, np.floor(w / w.sum() * ntrials_total)).astype(np.int64)
    counts[-1] += ntrials_total - counts.sum()
    subj = np.repeat(np.arange(1, nsubs + 1) * 3 + 100, counts)  # non-contiguous ids, like the CSV
    pe = np.clip(rng.normal(0.0, 5.8, ntrials_total), -35, 35)
    return subj, pe
