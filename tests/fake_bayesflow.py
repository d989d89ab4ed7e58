"""A stand-in for the slice of BayesFlow 1.1's simulation API that the reference uses
(basic_ddm_dc.py:130-134; bayesflow==1.1.3.dev4, yaml/bayesflow.yml:37) -- TEST INFRASTRUCTURE ONLY.

BayesFlow and TensorFlow are not installed in this image.  This stub reproduces the call
pattern of ``bf.simulation``: how Prior / ContextGenerator / Simulator / GenerativeModel invoke the
user's callables and which keys the resulting dict carries, so that the drop-in functions can be
exercised through the reference's own wiring code.  Semantics follow BayesFlow 1.1:
  * ``Prior(prior_fun=f)`` stacks ``f()`` batch_size times; ``Prior(batch_prior_fun=g)`` calls ``g(batch_size)``;
  * ``ContextGenerator(non_batchable_context_fun=h)`` calls ``h()`` once per batch;
  * ``Simulator(simulator_fun=s)`` calls ``s(params[b], non_batchable_context)`` for each b and stacks;
    ``Simulator(batch_simulator_fun=t)`` calls ``t(params, non_batchable_context)`` once;
  * ``GenerativeModel(prior, simulator)`` runs a 2-dataset self-test on construction and returns a dict
    with 'prior_draws', 'sim_data', 'sim_non_batchable_context' (+ the unused context keys).
"""
import types

import numpy as np


class ContextGenerator:
    def __init__(self, batchable_context_fun=None, non_batchable_context_fun=None, use_non_batchable_for_batchable=False):
        self.batchable_context_fun = batchable_context_fun
        self.non_batchable_context_fun = non_batchable_context_fun

    def __call__(self, batch_size):
        out = {"non_batchable_context": None, "batchable_context": None}
        if self.non_batchable_context_fun is not None:
            out["non_batchable_context"] = self.non_batchable_context_fun()
        if self.batchable_context_fun is not None:
            out["batchable_context"] = [self.batchable_context_fun() for _ in range(batch_size)]
        return out


class Prior:
    def __init__(self, batch_prior_fun=None, prior_fun=None, context_generator=None, param_names=None):
        if (batch_prior_fun is None) == (prior_fun is None):
            raise ValueError("Either batch_prior_fun or prior_fun should be provided, but not both!")
        self.prior = prior_fun if prior_fun is not None else batch_prior_fun
        self.is_batched = batch_prior_fun is not None
        self.param_names = param_names

    def __call__(self, batch_size, *args, **kwargs):
        if self.is_batched:
            draws = self.prior(batch_size, *args, **kwargs)
        else:
            draws = np.array([self.prior(*args, **kwargs) for _ in range(batch_size)])
        return {"prior_draws": draws, "batchable_context": None, "non_batchable_context": None}


class Simulator:
    def __init__(self, batch_simulator_fun=None, simulator_fun=None, context_generator=None):
        if (batch_simulator_fun is None) == (simulator_fun is None):
            raise ValueError("Either batch_simulator_fun or simulator_fun should be provided, but not both!")
        self.simulator = simulator_fun if simulator_fun is not None else batch_simulator_fun
        self.is_batched = batch_simulator_fun is not None
        self.context_gen = context_generator

    def __call__(self, params, *args, **kwargs):
        batch_size = params.shape[0]
        ctx = self.context_gen(batch_size) if self.context_gen is not None else {"non_batchable_context": None,
                                                                                   "batchable_context": None}
        nb = ctx["non_batchable_context"]
        extra = () if nb is None else (nb,)
        if self.is_batched:
            sim_data = self.simulator(params, *extra, *args, **kwargs)
        else:
            sim_data = np.array([self.simulator(params[b], *extra, *args, **kwargs) for b in range(batch_size)])
        return {"sim_data": sim_data, "batchable_context": ctx["batchable_context"], "non_batchable_context": nb}


class GenerativeModel:
    _N_SIM_TEST = 2

    def __init__(self, prior, simulator, skip_test=False, prior_is_batched=None, simulator_is_batched=None, name="anonymous"):
        self.prior, self.simulator, self.name = prior, simulator, name
        if not skip_test:
            out = self(self._N_SIM_TEST)
            assert out["prior_draws"].shape[0] == self._N_SIM_TEST and len(out["sim_data"]) == self._N_SIM_TEST

    def __call__(self, batch_size, **kwargs):
        p = self.prior(batch_size)
        s = self.simulator(p["prior_draws"])
        return {"prior_non_batchable_context": p["non_batchable_context"], "prior_batchable_context": p["batchable_context"],
                "prior_draws": p["prior_draws"], "sim_non_batchable_context": s["non_batchable_context"],
                "sim_batchable_context": s["batchable_context"], "sim_data": s["sim_data"]}


def as_module():
    """A module object shaped like ``bayesflow`` with the ``simulation`` sub-module."""
    bf = types.ModuleType("bayesflow")
    sim = types.ModuleType("bayesflow.simulation")
    sim.Prior, sim.ContextGenerator, sim.Simulator, sim.GenerativeModel = Prior, ContextGenerator, Simulator, GenerativeModel
    bf.simulation = sim
    return bf, sim
