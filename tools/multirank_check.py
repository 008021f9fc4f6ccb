"""Run under torchrun on N GPUs: landmark-sharded Global BA (sorted input -> device set-up path, shuffled input -> host sorting
path, LocalBA flags) against the CPU oracle on rank 0.  Prints one OK/FAIL line per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from ccm_slam_b200 import api, synth
from oracle import pyoracle as orc

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group(backend="gloo")
api.init(local)
uid = torch.from_numpy(api.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
dist.broadcast(uid, src=0)
api.comm_init(rank, world, uid.numpy())
orc.lib()


def close(a, b, tol):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def shard(p, r):
    import ctypes as C
    L0, L1, E0, E1 = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
    mp = np.ascontiguousarray(p.obs_mp, np.int32)
    rc = api.lib().ccm_ba_shard_range(mp.ctypes.data_as(C.c_void_p), p.E, p.P, r, world, C.byref(L0), C.byref(L1), C.byref(E0), C.byref(E1))
    assert rc == 0
    return L0.value, L1.value


def case(name, p, iters, delta):
    # every rank checks its own shard (landmarks [L0, L1) and their observations) against the oracle; poses are replicated
    ref = orc.ba_solve(p, iterations=iters, huber_delta=delta)
    res = api.ba_solve(p, iterations=iters, huber_delta=delta)
    L0, L1 = shard(p, rank)
    mine = (p.obs_mp >= L0) & (p.obs_mp < L1)
    act = mine if p.edge_flags is None else mine & ((p.edge_flags & 1) == 0)
    checks = dict(iters=res["iters_done"] == ref["iters_done"] and res["trials_total"] == ref["trials_total"],
                  trace=np.allclose(res["trace"][:len(ref["trace"]), 2], ref["trace"][:, 2], rtol=1e-7),
                  poses=close(res["poses"], ref["poses"], 1e-6), points=close(res["points"][L0:L1], ref["points"][L0:L1], 1e-6),
                  chi2=np.allclose(res["chi2"][act], ref["chi2"][act], rtol=1e-5, atol=1e-8),
                  depth=np.array_equal(res["depth_pos"][mine], ref["depth_pos"][mine]))
    ok = all(checks.values())
    print("%s rank %d %-16s N=%d shard [%d,%d) iters %d chi2 %.3f %s" % ("OK  " if ok else "FAIL", rank, name, world, L0, L1, res["iters_done"],
                                                                    res["chi2_final"], "" if ok else str(checks)), flush=True)
    dist.barrier()


# CCM_PCG_IMPL=2 in the environment sends these small systems through the streamed / distributed solve too (by size they take the
# replicated small-system kernel)
for cfg, it in (("small", 8), ("cfg3", 10), ("cfg4", 10)):
    p = synth.make_config(cfg)
    case(cfg + " sorted", p, it, api.HUBER_GBA)
p = synth.make_config("cfg2")
perm = np.random.default_rng(0).permutation(p.E)
q = p.copy(); q.obs_kf, q.obs_mp, q.obs_uv, q.obs_w = p.obs_kf[perm], p.obs_mp[perm], p.obs_uv[perm], p.obs_w[perm]
case("cfg2 shuffled", q, 8, api.HUBER_GBA)
q = p.copy(); q.edge_flags = (np.random.default_rng(1).random(p.E) < 0.1).astype(np.uint8) | 2
case("cfg2 flags", q, 8, api.HUBER_LOCAL)
api.comm_destroy()
