"""Multi-GPU plumbing on CPU: world_size-2 (and 3, ragged) gloo process groups run the shard /
all-gather logic the B200 path uses with NCCL; the gathered records must equal the
single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soccernet_calibration_sportlight_b200 import sharding
from tests import camera_inputs as CI, camera_parity as CP


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the per-frame work of this shard: the camera solve (host build of the kernel source)
        preds = CI.synthetic_predictions(n_frames, seed=77)
        lo, hi = sharding.shard_bounds(n_frames, rank, world)
        solve = CP.host_solver()
        local = torch.from_numpy(solve(preds[lo:hi], "original_voter", 0.5))
        full = sharding.all_gather_records(local, n_frames)
        if rank == 0:
            q.put(full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 16), (3, 10)])
def test_gloo_all_gather_matches_single_process(world, n_frames):
    CP.host_solver()                                     # build once, before forking
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = CP.host_solver()(CI.synthetic_predictions(n_frames, seed=77), "original_voter", 0.5)
    assert got.shape == (n_frames, 16)
    assert np.array_equal(got, single)


def test_no_process_group_is_identity():
    x = torch.arange(32.0).reshape(2, 16)
    assert sharding.all_gather_records(x, 2) is x
    with pytest.raises(ValueError):
        sharding.all_gather_records(x, 3)
