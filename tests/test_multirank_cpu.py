"""world_size-2 (and 4) CPU test of the N>1 host path over gloo: rendezvous, unique-id style broadcast, the landmark
shard cut every rank derives independently, and max-over-ranks timing — everything bench.py does around the GPU work."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ccm_slam_b200 import api, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _shard(p, rank, world):
    mpa = np.ascontiguousarray(p.obs_mp, np.int32)
    L0, L1 = C.c_int32(), C.c_int32(); E0, E1 = C.c_int64(), C.c_int64()
    rc = api.lib().ccm_ba_shard_range(mpa.ctypes.data_as(C.c_void_p), p.E, p.P, rank, world, C.byref(L0), C.byref(L1), C.byref(E0), C.byref(E1))
    assert rc == 0
    return L0.value, L1.value, E0.value, E1.value


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    p = synth.make_config("small")                      # every rank generates identical bytes from the seed
    uid = torch.from_numpy(np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, np.uint8))
    dist.broadcast(uid, src=0)                           # how the NCCL unique id travels in bench.py
    assert uid.numpy().tolist() == list(range(128))
    L0, L1, E0, E1 = _shard(p, rank, world)
    mine = torch.tensor([L0, L1, E0, E1], dtype=torch.int64)
    allr = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, mine)
    # the shards partition landmarks and observations, in rank order, balanced by observations
    assert allr[0][0] == 0 and allr[-1][1] == p.P and allr[0][2] == 0 and allr[-1][3] == p.E
    for a, b in zip(allr[:-1], allr[1:]):
        assert a[1] == b[0] and a[3] == b[2]
    sizes = np.array([int(a[3] - a[2]) for a in allr])
    assert sizes.max() - sizes.min() <= 2 * np.bincount(p.obs_mp).max()
    # a per-pose partial sum over the shard, all-reduced, equals the global sum (what the Hpp all-reduce relies on)
    sel = (p.obs_mp >= L0) & (p.obs_mp < L1)
    part = torch.from_numpy(np.bincount(p.obs_kf[sel], weights=p.obs_w[sel].astype(np.float64), minlength=p.K))
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    assert np.allclose(part.numpy(), np.bincount(p.obs_kf, weights=p.obs_w.astype(np.float64), minlength=p.K))
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)             # max-over-ranks timing
    assert abs(float(t[0]) - 0.1 * world) < 1e-12
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


@pytest.mark.parametrize("world", [2, 4])
def test_shard_partition_and_rendezvous_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs: pr.start()
    for pr in procs:
        pr.join(timeout=180)
        assert pr.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == list(range(world))
