"""Shared pieces of the model modules: configurators and the generative-model dict."""
from __future__ import annotations

import numpy as np

from . import _capi
from .simulator import DDMSimulator, default_simulator


def configurator(sim_dict):
    """numpy configurator with the reference's contract (basic_ddm_dc.py:139-160;
    single_trial_alpha_not_scaled.py:169-191 tolerates ``prior_draws=None``)."""
    out = dict()
    data = np.asarray(sim_dict['sim_data']).astype(np.float32)
    out['summary_conditions'] = data
    N = np.log(sim_dict['sim_non_batchable_context'])
    out['direct_conditions'] = N * np.ones((data.shape[0], 1), dtype=np.float32)
    if sim_dict.get('prior_draws', None) is not None:
        out['parameters'] = np.asarray(sim_dict['prior_draws']).astype(np.float32)
    return out


def device_configurator(sim_dict):
    """Same contract, device-resident: ``sim_data`` is a DeviceBatch (DLPack producer) or a
    torch tensor; returns torch tensors on that device (float32).  SURVEY.md section 8f-2."""
    import torch

    data = sim_dict['sim_data']
    if not isinstance(data, torch.Tensor):
        data = torch.from_dlpack(data)
    data = data.to(torch.float32)
    out = {'summary_conditions': data}
    n = float(np.log(sim_dict['sim_non_batchable_context']))
    out['direct_conditions'] = torch.full((data.shape[0], 1), n, dtype=torch.float32, device=data.device)
    if sim_dict.get('prior_draws', None) is not None:
        out['parameters'] = torch.as_tensor(np.asarray(sim_dict['prior_draws'], dtype=np.float32), device=data.device)
    return out


class ModelAPI:
    """Builds the reference-shaped callables of one model variant."""

    def __init__(self, model_id: int, prior_name: str, dt: float = 0.01, max_steps: float = 400., flags: int = 0):
        self.model_id = model_id
        self.prior_name = prior_name
        self.n_params = _capi.N_PARAMS[model_id]
        self.dt = dt
        self.max_steps = max_steps
        self.flags = flags

    def sim(self, simulator: DDMSimulator | None) -> DDMSimulator:
        return simulator if simulator is not None else default_simulator()

    def simulate_trials(self, params, n_trials, simulator=None, **kw):
        """``simulator_fun`` contract: (P,) f64, int -> (n_trials, 2) f64."""
        params = np.asarray(params, dtype=np.float64).ravel()
        if params.size != self.n_params:
            raise ValueError(f"expected {self.n_params} parameters, got {params.size}")
        return self.batch_simulate_trials(params[None, :], n_trials, simulator, **kw)[0]

    def batch_simulate_trials(self, params, n_trials, simulator=None, *, dt=None, max_steps=None, precision=32,
                              flags=None, seed=None, dataset_offset=None, out=None):
        """``batch_simulator_fun`` contract: (B, P) f64, int -> (B, n_trials, 2) f64."""
        return self.sim(simulator).simulate(
            self.model_id, params, int(n_trials), self.dt if dt is None else dt,
            int(self.max_steps if max_steps is None else max_steps), precision=precision,
            flags=self.flags if flags is None else flags, seed=seed, dataset_offset=dataset_offset, out=out)

    def batch_simulate_histogram(self, params, n_trials, simulator=None, *, dt=None, max_steps=None, precision=32,
                                 flags=None, seed=None, dataset_offset=None, n_bins=400, rt_max=4.0):
        """(B, P) f64 host parameters -> the batch's response-time histogram by boundary, reduced on the device
        ("outputs reduced on device"); only 2 n_bins + 2 counters come back."""
        return self.sim(simulator).simulate_histogram(
            self.model_id, params, int(n_trials), self.dt if dt is None else dt,
            int(self.max_steps if max_steps is None else max_steps), precision=precision,
            flags=self.flags if flags is None else flags, seed=seed, dataset_offset=dataset_offset, n_bins=n_bins,
            rt_max=rt_max)

    def batch_simulate_trials_device(self, params, n_trials, simulator=None, *, dt=None, max_steps=None,
                                     precision=32, flags=None, seed=None, dataset_offset=None):
        """Device-resident variant: returns a DLPack producer of shape (B, n_trials, 2) float32."""
        f = (self.flags if flags is None else flags) | _capi.FLAG_OUT_F32
        return self.sim(simulator).simulate_device(
            self.model_id, params, int(n_trials), self.dt if dt is None else dt,
            int(self.max_steps if max_steps is None else max_steps), precision=precision, flags=f, seed=seed,
            dataset_offset=dataset_offset)

    def generative_model(self, batch_size, draw_batch, prior_N, simulator=None, device=False, device_prior=False):
        """What ``bf.simulation.GenerativeModel(prior, simulator)(batch_size)`` returns to the
        reference's configurator: dict with prior_draws, sim_data, sim_non_batchable_context.
        ``device_prior=True`` draws the parameters on the GPU too (``ddm_draw_prior``): prior and
        simulation are two launches and the parameters never cross PCIe on their way in."""
        n = int(prior_N())
        if device_prior and device:  # the whole batch in one call: one FFI crossing, one stream synchronisation
            prior_draws, data = self.sim(simulator).training_batch(self.prior_name, batch_size, n, self.dt, int(self.max_steps),
                                                                   flags=self.flags | _capi.FLAG_OUT_F32)
            return {'prior_draws': prior_draws, 'sim_data': data, 'sim_non_batchable_context': n}
        if device_prior:
            sim = self.sim(simulator)
            prior_draws = sim.draw_prior(self.prior_name, batch_size)
            f = self.flags | (_capi.FLAG_OUT_F32 if device else 0)
            sim.run_uploaded(n, self.dt, int(self.max_steps), flags=f)
            data = sim.last_output_dlpack() if device else sim.download((int(batch_size), n, 2), False)
            return {'prior_draws': prior_draws, 'sim_data': data, 'sim_non_batchable_context': n}
        prior_draws = draw_batch(batch_size)
        if device:
            data = self.batch_simulate_trials_device(prior_draws, n, simulator)
        else:
            data = self.batch_simulate_trials(prior_draws, n, simulator)
        return {'prior_draws': prior_draws, 'sim_data': data, 'sim_non_batchable_context': n}


def bayesflow_generative_model(draw_prior, prior_N, simulate_trials=None, batch_simulate_trials=None,
                               batch_prior=None):
    """The reference's wiring (basic_ddm_dc.py:130-134) on BayesFlow 1.1, if it is installed.
    With ``batch_simulate_trials`` the whole batch is one kernel launch."""
    import bayesflow as bf  # noqa: not installed in the build container

    prior = (bf.simulation.Prior(batch_prior_fun=batch_prior) if batch_prior is not None
             else bf.simulation.Prior(prior_fun=draw_prior))
    context = bf.simulation.ContextGenerator(non_batchable_context_fun=prior_N)
    if batch_simulate_trials is not None:
        simulator = bf.simulation.Simulator(batch_simulator_fun=batch_simulate_trials, context_generator=context)
    else:
        simulator = bf.simulation.Simulator(simulator_fun=simulate_trials, context_generator=context)
    return bf.simulation.GenerativeModel(prior, simulator)
