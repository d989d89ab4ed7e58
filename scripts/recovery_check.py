"""Proxy for correctness check #3 (posterior recovery) without BayesFlow/TensorFlow (absent from this image).

An amortized point estimator (DeepSets summary network -> posterior-mean regression, PyTorch) is trained
online on batches from the CUDA simulator (device prior -> kernel -> DLPack -> torch, nothing through the
host), exactly the data path BayesFlow's trainer would use.  It is then evaluated on two held-out test sets
built from the SAME prior draws: one simulated by the CUDA kernel, one by the CPU oracle of the reference's
numba loop on its own MT19937 stream.  If the two simulators are interchangeable for inference, recovery
(R^2 and Pearson r of posterior mean vs truth per parameter -- the reference's own recovery metric,
pyhddmjagsutils.py:609-623) must agree between the two test sets.

    python scripts/recovery_check.py [--iters 4000] [--out profiles/r01_recovery_check.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesflow_nddms_b200 as pkg  # noqa: E402
from bayesflow_nddms_b200 import basic_ddm_dc as m  # noqa: E402
from oracle import cpu as orc  # noqa: E402

NAMES = ["drift", "alpha", "beta", "ter", "dc"]


class Estimator(nn.Module):
    def __init__(self, d_in=2, width=128, n_params=5):
        super().__init__()
        self.phi = nn.Sequential(nn.Linear(d_in, width), nn.SiLU(), nn.Linear(width, width), nn.SiLU(), nn.Linear(width, width))
        self.rho = nn.Sequential(nn.Linear(2 * width + 1, width), nn.SiLU(), nn.Linear(width, width), nn.SiLU(),
                                 nn.Linear(width, n_params))

    def forward(self, x, log_n):
        h = self.phi(x)
        s = torch.cat([h.mean(1), h.amax(1), log_n], dim=1)
        return self.rho(s)


def recovery(truth, est):
    out = {}
    for j, name in enumerate(NAMES):
        t, e = truth[:, j], est[:, j]
        r2 = 1.0 - np.sum((t - e) ** 2) / np.sum((t - t.mean()) ** 2)
        out[name] = {"r2": float(r2), "pearson": float(np.corrcoef(t, e)[0, 1])}
    return out


def run(iters=4000, batch=64, n_test=1000, n_trials_test=200, seed=0, device=0, verbose=True):
    torch.manual_seed(seed)
    np.random.seed(seed)
    dev = torch.device("cuda", device)
    sim = pkg.DDMSimulator(device=device, seed=1234)
    net = Estimator().to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, iters)
    mean = torch.tensor([0.0, 1.0, 0.5, 0.5, 1.0], device=dev)
    std = torch.tensor([2.0, 0.5, 0.22, 0.25, 0.5], device=dev)
    t0 = time.time()
    for it in range(iters):
        d = m.generative_model(batch, sim, device=True, device_prior=True)     # N ~ U{60..300}, as the reference
        c = m.device_configurator(d)
        x, y, ln = c['summary_conditions'], c['parameters'], c['direct_conditions']
        loss = ((net(x, ln) - (y - mean) / std) ** 2).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        sched.step()
        if verbose and (it % 500 == 0 or it == iters - 1):
            print(f"iter {it:5d} loss {loss.item():.4f}  ({time.time() - t0:.0f} s)", flush=True)
    train_s = time.time() - t0
    # ---- held-out test sets from the same prior draws ----
    rng = np.random.default_rng(99)
    from bayesflow_nddms_b200 import priors

    theta = priors.draw_prior_batch("basic", n_test, rng)
    gpu_data = sim.simulate(0, theta, n_trials_test, seed=777, dataset_offset=10_000_000)
    cpu_data, _, _ = orc.simulate_batch_mt(0, theta, n_trials_test, seed=4242, n_threads=max(1, os.cpu_count() or 1))
    net.eval()
    res = {}
    with torch.no_grad():
        ln = torch.full((n_test, 1), float(np.log(n_trials_test)), device=dev)
        for name, data in (("cuda_simulator", gpu_data), ("reference_loop_cpu", cpu_data)):
            est = net(torch.as_tensor(data, dtype=torch.float32, device=dev), ln) * std + mean
            res[name] = recovery(theta, est.cpu().numpy().astype(np.float64))
    diff = {p: {"d_r2": res["cuda_simulator"][p]["r2"] - res["reference_loop_cpu"][p]["r2"],
                "d_pearson": res["cuda_simulator"][p]["pearson"] - res["reference_loop_cpu"][p]["pearson"]} for p in NAMES}
    sim.close()
    return {"what": "amortized posterior-mean estimator trained on the CUDA simulator; recovery on held-out data from the CUDA "
                    "simulator vs from the CPU oracle of the reference loop (same prior draws)",
            "train_iterations": iters, "batch_size": batch, "train_seconds": train_s, "n_test_datasets": n_test,
            "n_trials_test": n_trials_test, "recovery": res, "difference": diff}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=4000)
    ap.add_argument("--out", default="gpurun_out/recovery_check.json")
    a = ap.parse_args()
    r = run(a.iters)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(r, open(a.out, "w"), indent=1)
    print(json.dumps(r["recovery"], indent=1))
    print(json.dumps(r["difference"], indent=1))
