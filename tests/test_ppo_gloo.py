"""World-size-2 (gloo, CPU) check of the trainer's data-parallel path: each rank trains on its own env shard, the flat
gradient buffer is averaged with one all-reduce per minibatch, and the replicas must stay bit-identical; the averaged
update must equal what one process computes from both shards' minibatch gradients."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
HERE = Path(__file__).resolve().parent


def _setup():
    for p in (str(ROOT), str(HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import _nav3d_path  # noqa: F401


def _worker(rank, world, port, out):
    _setup()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from nav3d.ppo import RecurrentPPO
    from test_ppo_host import tiny_rooms
    from train_refs import OracleBatchedEnv, TorchOps
    env = OracleBatchedEnv(tiny_rooms(), 4, seed=100 + rank)           # different data per rank
    # different torch seeds per rank: the broadcast from rank 0 must make the replicas identical anyway
    m = RecurrentPPO(env, policy_kwargs=dict(net_arch=dict(pi=[16], vf=[16]), lstm_hidden_size=16), n_steps=32, batch_size=64,
                     seq_len=16, n_epochs=2, seed=rank, ops=TorchOps(rank))
    assert m.world == 2 and m.rank == rank
    m.learn(total_timesteps=2 * 32 * 4 * world)
    assert m.num_timesteps == 2 * 32 * 4 * world                        # timesteps count all ranks' envs
    flat = torch.cat([p.detach().flatten() for p in m.policy.parameters()])
    parts = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    if rank == 0:
        np.savez(out, a=parts[0].numpy(), b=parts[1].numpy(), rew=np.array([r["rollout_reward_mean"] for r in m.logger]))
    dist.barrier()
    dist.destroy_process_group()


def test_replicas_stay_identical(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "params.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    assert np.array_equal(z["a"], z["b"]) and np.isfinite(z["a"]).all()
    assert len(z["rew"]) == 2


def test_flat_gradient_average_equals_two_shard_mean():
    """One process, two 'ranks' by hand: averaging the flat gradient buffers equals the gradient of the mean loss."""
    _setup()
    from nav3d.policy import RecurrentActorCritic
    torch.manual_seed(0)
    pol = RecurrentActorCritic(obs_dim=8, n_actions=3, net_arch=dict(pi=[8], vf=[8]), lstm_hidden_size=8).double()
    xs = [torch.randn(5, 4, 8, dtype=torch.float64) for _ in range(2)]
    st = [torch.zeros(5, 4, dtype=torch.uint8) for _ in range(2)]

    def loss_of(x, s):
        l, v, _ = pol.forward_sequence(x, pol.initial_state(4, dtype=torch.float64), s)
        return l.square().mean() + v.square().mean()
    grads = []
    for x, s in zip(xs, st):
        pol.zero_grad()
        loss_of(x, s).backward()
        grads.append(torch.cat([p.grad.flatten() for p in pol.parameters()]))
    pol.zero_grad()
    (0.5 * (loss_of(xs[0], st[0]) + loss_of(xs[1], st[1]))).backward()
    both = torch.cat([p.grad.flatten() for p in pol.parameters()])
    assert torch.allclose(0.5 * (grads[0] + grads[1]), both, atol=1e-12)
