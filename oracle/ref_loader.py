"""Load the reference's VERBATIM hot-path functions by line range -- TEST INFRASTRUCTURE ONLY.

The reference scripts cannot be imported whole (``import bayesflow`` at
basic_ddm_dc.py:29 and module-level network construction), but the hot-path
functions are self-contained.  This module reads source line ranges from
``/root/reference`` (present only in the build container), dedents them and
exec's them with ``{np, njit, truncnorm}`` in scope.  No reference source is
copied into this repository: the text is read at run time, used to generate
``tests/golden/*.npz`` (see tests/golden/make_golden.py), and discarded.

Nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.
"""
from __future__ import annotations

import os
import textwrap

REFERENCE_ROOT = os.environ.get("DDM_REFERENCE_ROOT", "/root/reference")

# (file, first line, last line) -- 1-based inclusive, as cited in SURVEY.md section 8a.
RANGES = {
    "basic": [("basic_ddm_dc.py", 85, 125)],
    "alpha": [("single_trial_alpha_not_scaled.py", 107, 155)],
    "alpha_dc": [("single_trial_alpha_not_scaled.py", 926, 974)],
    "alpha_scale": [("single_trial_alpha_not_scaled.py", 1237, 1285)],
    "alpha_scale2": [("single_trial_alpha_not_scaled.py", 1471, 1519)],
    # simulate_trials_fine calls the M1 diffusion_trial with dt=.001, max_steps=4000
    "alpha_fine": [("single_trial_alpha_not_scaled.py", 107, 155),
                   ("single_trial_alpha_not_scaled.py", 1710, 1722)],
    "stahl": [("imputation_from_stahl_not_scaled.py", 120, 148)],
    "eta": [("retired_models/basic_ddm_eta_dc.py", 80, 120)],
    "drift_dc5": [("retired_models/single_trial_drift_dc5.py", 90, 154)],
    "drift_dc4": [("retired_models/single_trial_drift_dc4.py", 90, 146)],
    "alpha_dc2ch": [("retired_models/single_trial_alpha_dc.py", 109, 176)],
    "drift_alpha": [("retired_models/single_trial_drift_alpha.py", 96, 152)],
    "alpha_std1": [("retired_models/single_trial_alpha.py", 83, 135)],
    "evidence": [("retired_models/basic_ddm_dc_evidence.py", 87, 151)],
    "evidence2": [("retired_models/basic_ddm_dc_evidence2.py", 83, 150)],
    "evidence_no_noise2": [("retired_models/basic_ddm_dc_evidence_no_noise2.py", 82, 147)],
    "basic_prior": [("basic_ddm_dc.py", 50, 81)],
    "alpha_prior": [("single_trial_alpha_not_scaled.py", 66, 103)],
}

ENTRY = {
    "basic": "simulate_trials",
    "alpha": "simulate_trials",
    "alpha_dc": "simulate_trials_alt",
    "alpha_scale": "simulate_trials_scale",
    "alpha_scale2": "simulate_trials_scale2",
    "alpha_fine": "simulate_trials_fine",
    "stahl": "diffusion_trial",
    "eta": "simulate_trials",
    "drift_dc5": "simulate_trials",
    "drift_dc4": "simulate_trials",
    "alpha_dc2ch": "simulate_trials",
    "drift_alpha": "simulate_trials",
    "alpha_std1": "simulate_trials",
    "evidence": "simulate_trials",
    "evidence2": "simulate_trials",
    "evidence_no_noise2": "simulate_trials",
}


def available() -> bool:
    return os.path.isdir(REFERENCE_ROOT)


def _read(file: str, first: int, last: int) -> str:
    with open(os.path.join(REFERENCE_ROOT, file), "r") as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def load(name: str) -> dict:
    """Exec the verbatim reference source for ``name``; return its namespace."""
    import numpy as np
    from numba import njit
    from scipy.stats import truncnorm

    ns = {"np": np, "njit": njit, "truncnorm": truncnorm}
    for file, first, last in RANGES[name]:
        code = compile(_read(file, first, last), f"{REFERENCE_ROOT}/{file}:{first}-{last}", "exec")
        exec(code, ns)
    return ns


def seeded_call(name: str, seed: int, *args):
    """Run the verbatim reference entry point after np.random.seed(seed) INSIDE
    jitted code (a Python-level seed does not reach numba's generator)."""
    import numpy as np
    from numba import njit

    ns = load(name)
    fn = ns[ENTRY[name]]

    @njit
    def _seed(s):
        np.random.seed(s)

    _seed(seed)
    return fn(*args)
