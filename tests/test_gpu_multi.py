"""Multi-GPU path on real devices: dataset sharding over ranks + the optional NCCL all-gather.
Needs >= 2 GPUs (skipped otherwise); the host-side logic is covered on CPU by the gloo test."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["LOCAL_RANK"] = str(rank)
    import torch
    import torch.distributed as dist

    import bayesflow_nddms_b200 as pkg
    from bayesflow_nddms_b200 import basic_ddm_dc as m
    from bayesflow_nddms_b200 import distributed as D
    from bayesflow_nddms_b200 import priors

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    sim = pkg.DDMSimulator(device=rank, seed=77)
    params = priors.draw_prior_batch("basic", 65, np.random.default_rng(3))   # 65: uneven shards
    full, (lo, hi) = D.simulate_sharded(lambda p, n, dataset_offset: m.batch_simulate_trials_device(
        p, n, sim, dataset_offset=dataset_offset), params, 300, dataset_base=1000, gather=True)
    even, _ = D.simulate_sharded(lambda p, n, dataset_offset: m.batch_simulate_trials_device(
        p, n, sim, dataset_offset=dataset_offset), params[:64], 300, dataset_base=1000, gather=True)
    assert full.is_cuda and full.device.index == rank
    q.put((rank, lo, hi, full.cpu().numpy(), even.cpu().numpy()))
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


def test_sharded_simulation_nccl_allgather(sim):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from bayesflow_nddms_b200 import priors

    params = priors.draw_prior_batch("basic", 65, np.random.default_rng(3))
    single = sim.simulate(0, params, 300, seed=77, dataset_offset=1000, flags=2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, lo, hi, full, even in res:
        assert np.array_equal(full, single)          # bit-identical to the one-GPU batch on every rank
        assert np.array_equal(even, single[:64])
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 65
