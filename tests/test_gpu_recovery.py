"""Proxy for check #3 (posterior recovery): an amortized estimator trained on CUDA-simulated batches recovers
parameters equally well from CUDA-simulated and from reference-loop (CPU oracle) test data.  BayesFlow and
TensorFlow are not installed here, so the networks are a small PyTorch stand-in (scripts/recovery_check.py)."""
import os
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600, method="thread")
def test_recovery_agrees_between_cuda_simulator_and_reference_loop():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import recovery_check

    r = recovery_check.run(iters=1500, batch=64, n_test=600, n_trials_test=200, verbose=False)
    rec, diff = r["recovery"], r["difference"]
    # the well-identified parameters are recovered (short training: loose floors) ...
    for p, floor in (("drift", 0.80), ("ter", 0.80), ("beta", 0.40)):
        assert rec["cuda_simulator"][p]["r2"] > floor and rec["reference_loop_cpu"][p]["r2"] > floor, (p, rec)
    # ... and equally well from either simulator's data (tolerance stated here: 0.06 in R^2, 0.04 in Pearson r)
    for p in recovery_check.NAMES:
        assert abs(diff[p]["d_r2"]) < 0.06 and abs(diff[p]["d_pearson"]) < 0.04, (p, diff[p], rec)
