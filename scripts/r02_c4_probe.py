"""Where the 0.12 ms of a C4-sized ddm_simulate_trialwise call go (kernel: 0.018 ms)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import _capi as capi
from bayesflow_nddms_b200 import imputation_from_stahl_not_scaled as stahl
sim = pkg.DDMSimulator(0, seed=2023)
subj, pe = stahl.synthetic_stahl_like()
pp = stahl.draw_participant_params(89, np.random.default_rng(2024))
_, alphas = stahl.boundaries_from_pe(pe)
_, idx = np.unique(subj, return_inverse=True)
idx = idx.astype(np.int32)
def med(f, reps=200):
    for _ in range(10): f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); sim.synchronize(); ts.append((time.perf_counter() - t0) * 1e6)
    return np.median(ts)
print("pageable in, host out      %.1f us" % med(lambda: sim.simulate_trialwise(idx, alphas, pp)))
pi = sim.pinned_empty(idx.shape, np.int32, "i"); pi[:] = idx
pa = sim.pinned_empty(alphas.shape, np.float64, "a"); pa[:] = alphas
ppp = sim.pinned_empty(pp.shape, np.float64, "p"); ppp[:] = pp
print("pinned in, host out        %.1f us" % med(lambda: sim.simulate_trialwise(pi, pa, ppp)))
def dev():
    b = sim.simulate_trialwise(pi, pa, ppp, flags=capi.FLAG_OUT_F32, device=True); del b
print("pinned in, device out      %.1f us" % med(dev))
def dev2():
    b = sim.simulate_trialwise(idx, alphas, pp, flags=capi.FLAG_OUT_F32, device=True); del b
print("pageable in, device out    %.1f us" % med(dev2))
print("kernel %.1f us" % (sim.last_stats()["kernel_ms"] * 1e3))
