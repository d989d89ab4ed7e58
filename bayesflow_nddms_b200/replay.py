"""Device-resident replay feed.

The reference trains with ``trainer.train_experience_replay(epochs, batch_size, iterations_per_epoch)``
(basic_ddm_dc.py:199-202): every iteration simulates a fresh batch, stores it in a FIFO memory and
trains on a batch sampled from the memory.  BayesFlow keeps that memory as host numpy dicts, so every
step pays a host->device copy.  ``DeviceReplayBuffer`` keeps configured batches as torch tensors on
the simulator's GPU (they arrive there through DLPack and never leave), with the same store / sample
behaviour: capacity counted in simulated batches, oldest overwritten first, one stored batch drawn
uniformly per sample (the number of trials N is shared by a batch, so batches are not mixed).
"""
from __future__ import annotations

import numpy as np


class DeviceReplayBuffer:
    def __init__(self, capacity_in_batches: int = 1000, rng=None):
        if capacity_in_batches < 1:
            raise ValueError("capacity must be >= 1")
        self.capacity = int(capacity_in_batches)
        self._slots = []
        self._next = 0
        self._rng = np.random.default_rng() if rng is None else rng
        self.stored_total = 0

    def __len__(self):
        return len(self._slots)

    def is_full(self):
        return len(self._slots) == self.capacity

    def store(self, configured: dict):
        """Keep one configured batch (dict of tensors/arrays as returned by ``device_configurator``)."""
        if len(self._slots) < self.capacity:
            self._slots.append(configured)
        else:
            self._slots[self._next] = configured
        self._next = (self._next + 1) % self.capacity
        self.stored_total += 1

    def sample(self) -> dict:
        if not self._slots:
            raise RuntimeError("replay buffer is empty")
        return self._slots[int(self._rng.integers(len(self._slots)))]

    def nbytes(self) -> int:
        n = 0
        for d in self._slots:
            for v in d.values():
                n += int(v.numel() * v.element_size()) if hasattr(v, "numel") else int(np.asarray(v).nbytes)
        return n


def replay_iterations(model_module, batch_size: int, n_iterations: int, buffer: DeviceReplayBuffer, simulator=None,
                      device_prior: bool = True):
    """Generator for the data side of ``train_experience_replay``: per iteration simulate a fresh batch on
    the GPU (prior + trials), configure it on the device, store it, and yield a sampled stored batch."""
    for _ in range(int(n_iterations)):
        d = model_module.generative_model(batch_size, simulator, device=True, device_prior=device_prior)
        buffer.store(model_module.device_configurator(d))
        yield buffer.sample()
