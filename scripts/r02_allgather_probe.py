"""2 GPUs: why does simulate + all-gather cost more than the sum of its parts?  torchrun --nproc-per-node 2 scripts/r02_allgather_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

import bayesflow_nddms_b200 as pkg
from bayesflow_nddms_b200 import distributed as D
from bayesflow_nddms_b200 import priors
from bayesflow_nddms_b200 import single_trial_alpha_not_scaled as m1

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
sim = pkg.DDMSimulator(device=rank, seed=2023)
B, N = 1024, 1000
P = priors.draw_prior_batch("alpha", B, np.random.default_rng(77))
lo, hi = D.shard_range(B, rank, world)
local = torch.from_dlpack(m1.batch_simulate_trials_device(P[lo:hi], N, sim, dataset_offset=5000 + lo))
full = D.all_gather_batch(local, B, world)
persistent = torch.empty_like(local)


def timeit(label, fn, reps=50):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps * 1e6
    if rank == 0:
        print(f"{label:60s} {dt:9.1f} us", flush=True)


state = {"local": local}


def gather_only():
    D.all_gather_batch(state["local"], B, world)


def sim_only():
    state["local"] = torch.from_dlpack(m1.batch_simulate_trials_device(P[lo:hi], N, sim, dataset_offset=5000 + lo))


def sim_gather():
    sim_only()
    D.all_gather_batch(state["local"], B, world)


def sim_copy_gather():
    sim_only()
    persistent.copy_(state["local"])
    D.all_gather_batch(persistent, B, world)


def sim_gather_sync():
    sim_only()
    D.all_gather_batch(state["local"], B, world)
    torch.cuda.synchronize()


timeit("all-gather only (fixed input)", gather_only)
timeit("simulate only (DLPack -> torch)", sim_only)
timeit("simulate + all-gather of the fresh DLPack tensor", sim_gather)
timeit("simulate + copy into a torch buffer + all-gather", sim_copy_gather)
timeit("simulate + all-gather + device sync", sim_gather_sync)
os.environ["X"] = "1"
sim.close()
dist.destroy_process_group()
