"""World-size-2 check (gloo, CPU) of the multi-GPU host logic: each rank derives its env shard (env_id0, n_local) the way
bench.py does, plays its envs with the oracle, and the gathered result must equal the single-process run — i.e. the
Philox streams are keyed by GLOBAL env id, so results do not depend on the number of shards.  (The data path itself has no
collective; gloo only carries the comparison.)"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def shard(rank, world, n_total):
    n_local = n_total // world
    return rank * n_local, n_local


def _play(env_id0, n, seed, T):
    sys.path.insert(0, str(ROOT))
    import _nav3d_path  # noqa: F401
    from nav3d.rooms import load_room_dir
    from oracle import c_oracle
    rooms = load_room_dir(ROOT / "rooms" / "P1_training", sort=True)
    ov = c_oracle.OracleVec(n, [c_oracle.OracleRoom(r.grid, -2) for r in rooms], 10, -2.0, seed, env_id0, True)
    obs0 = ov.reset().copy()
    rs = []
    for t in range(T):
        a = np.array([c_oracle.action(seed, env_id0 + i, t) for i in range(n)])
        ov.step(a)
        rs.append(ov.reward.copy())
    return obs0, ov.obs.copy(), np.stack(rs), ov.state()[:, :14]


def _worker(rank, world, port, n_total, seed, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env_id0, n_local = shard(rank, world, n_total)
    obs0, obs, rew, st = _play(env_id0, n_local, seed, T)
    parts = [torch.zeros((n_local, 80)) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(obs))
    rparts = [torch.zeros((T, n_local), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(rparts, torch.from_numpy(rew))
    sparts = [torch.zeros((n_local, 14), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sparts, torch.from_numpy(st))
    tmax = torch.tensor([float(rank + 1)])
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)           # the max-over-ranks timing reduction of bench.py
    assert tmax.item() == world
    if rank == 0:
        np.savez(out, obs=torch.cat(parts).numpy(), rew=torch.cat(rparts, dim=1).numpy(), st=torch.cat(sparts).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_equal_one(tmp_path):
    n_total, seed, T = 64, 5, 40
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, n_total, seed, T, out), nprocs=2, join=True)
    got = np.load(out)
    _, obs, rew, st = _play(0, n_total, seed, T)
    assert np.array_equal(got["obs"].view(np.uint32), obs.view(np.uint32))
    assert np.array_equal(got["rew"], rew)
    assert np.array_equal(got["st"], st)


def test_shard_arithmetic():
    assert [shard(r, 8, 1 << 20) for r in (0, 7)] == [(0, 131072), (917504, 131072)]
    assert shard(1, 2, 1 << 20) == (524288, 524288)
