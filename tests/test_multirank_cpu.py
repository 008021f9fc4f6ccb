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


def _shard_problem(p, L0, L1):
    """the BA problem restricted to landmarks [L0, L1): every pose, the shard's points and their observations"""
    sel = (p.obs_mp >= L0) & (p.obs_mp < L1)
    return synth.BAProblem(poses=p.poses, intr=p.intr, fixed=p.fixed, points=p.points[L0:L1], obs_kf=p.obs_kf[sel],
                           obs_mp=(p.obs_mp[sel] - L0).astype(np.int32), obs_uv=p.obs_uv[sel], obs_w=p.obs_w[sel])


def _worker_reduced_system(rank, world, port, q):
    """what the N>1 BA path rests on: the reduced camera system is a sum over landmarks, so each rank builds the contribution of its own
    landmark shard and one all-reduce yields the system every rank then solves — here with the CPU oracle's pieces over gloo"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from oracle import pyoracle as orc
    p = synth.make_config("small")
    lam = 3.7
    L0, L1, E0, E1 = _shard(p, rank, world)
    mine = orc.ba_schur_solve(_shard_problem(p, L0, L1), lam, dense=True)
    S = torch.from_numpy(mine["S"].copy()); b = torch.from_numpy(mine["bschur"].copy())
    dist.all_reduce(S, op=dist.ReduceOp.SUM); dist.all_reduce(b, op=dist.ReduceOp.SUM)
    S = S.numpy(); b = b.numpy()
    free = np.repeat(p.fixed == 0, 6)
    sel = (p.obs_mp >= L0) & (p.obs_mp < L1)
    seen = torch.from_numpy((np.bincount(p.obs_kf[sel], minlength=p.K) > 0).astype(np.float64))    # poses this shard's optimiser knows
    dist.all_reduce(seen, op=dist.ReduceOp.SUM)
    S[np.diag_indices_from(S)] -= (np.repeat(seen.numpy(), 6) - 1) * lam * free   # every shard that sees a pose damped its diagonal once
    whole = orc.ba_schur_solve(p, lam, dense=True)
    assert np.allclose(S, whole["S"], rtol=1e-11, atol=1e-9 * np.abs(whole["S"]).max())
    assert np.allclose(b, whole["bschur"], rtol=1e-11, atol=1e-9 * np.abs(whole["bschur"]).max())
    # every rank solves the same system and back-substitutes its own landmarks: the shard's point updates are the global ones
    idx = np.flatnonzero(free)
    dxp = np.zeros(6 * p.K); dxp[idx] = np.linalg.solve(S[np.ix_(idx, idx)], b[idx])
    assert np.allclose(dxp.reshape(p.K, 6), whole["dx_pose"], rtol=1e-7, atol=1e-9 * np.abs(whole["dx_pose"]).max())
    built = orc.ba_build(_shard_problem(p, L0, L1))
    sp = _shard_problem(p, L0, L1)
    dxl = np.zeros((L1 - L0, 3))
    rhs = built["bl"].copy()
    for e in range(sp.E):
        rhs[sp.obs_mp[e]] -= built["W"][e].T @ dxp[6 * sp.obs_kf[e]:6 * sp.obs_kf[e] + 6]
    for l in range(L1 - L0):
        dxl[l] = np.linalg.solve(built["Hll"][l] + lam * np.eye(3), rhs[l])
    assert np.allclose(dxl, whole["dx_point"][L0:L1], rtol=1e-6, atol=1e-9 * np.abs(whole["dx_point"]).max())
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


@pytest.mark.parametrize("world", [2, 3])
def test_landmark_sharded_reduced_system_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_reduced_system, args=(r, world, port, q)) for r in range(world)]
    for pr in procs: pr.start()
    for pr in procs:
        pr.join(timeout=240)
        assert pr.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == list(range(world))
