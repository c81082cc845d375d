"""GPU parity of the secondary env (NAV3D_ENV_SIMPLE == the reference's envs/simpleEnv.py) through the C ABI."""
import random

import numpy as np
import pytest
import torch

from conftest import ROOMS
from lockstep import replay_simple_golden_trace
from nav3d.rooms import load_room_dir, load_room_file

pytestmark = pytest.mark.gpu


class GpuSimple:
    def __init__(self, n, rooms, L, seed=0, env_id0=0, auto_reset=False, lanes=0):
        from nav3d import Engine, _lib
        self.n = n
        self.eng = Engine(n, rooms, local_map_length=L, auto_reset=auto_reset, seed=seed, env_id0=env_id0,
                          lanes_per_env=lanes, env_kind=_lib.ENV_SIMPLE)
        dev, d = self.eng.device, self.eng.obs_dim
        self.d_obs = torch.full((n, d), float("nan"), device=dev)
        self.d_rew = torch.zeros(n, device=dev); self.d_rew64 = torch.zeros(n, dtype=torch.float64, device=dev)
        self.d_te = torch.zeros(n, dtype=torch.uint8, device=dev); self.d_tr = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.d_tobs = torch.zeros((n, d), device=dev)

    def _pull(self):
        self.obs, self.reward, self.reward64 = self.d_obs.cpu().numpy(), self.d_rew.cpu().numpy(), self.d_rew64.cpu().numpy()
        self.term, self.trunc, self.tobs = self.d_te.cpu().numpy(), self.d_tr.cpu().numpy(), self.d_tobs.cpu().numpy()

    def reset(self, picks=None):
        self.eng.reset(self.d_obs, picks=None if picks is None else torch.as_tensor(np.asarray(picks, dtype=np.int32)))
        self._pull()
        return self.obs

    def step(self, actions):
        self.eng.step(torch.as_tensor(np.asarray(actions, dtype=np.int64)).to(self.eng.device), self.d_obs, self.d_rew, self.d_te,
                      self.d_tr, reward64=self.d_rew64, terminal_obs=self.d_tobs)
        self._pull()

    def state(self):
        return self.eng.get_state().cpu().numpy()

    def grid(self, env):
        return self.eng.get_grid(env)


@pytest.mark.parametrize("lanes", [1, 4, 32])
def test_simple_reference_golden_traces(simple_traces, lanes):
    for c in simple_traces:
        room = load_room_file(ROOMS / c["room"], simple=True)
        replay_simple_golden_trace(c, room, lambda rooms, L: GpuSimple(1, rooms, L, lanes=lanes))


@pytest.mark.parametrize("lanes,L", [(0, 4), (8, 10), (2, 1)])
def test_simple_lockstep_autoreset(oracle, lanes, L):
    rooms = [load_room_file(ROOMS / "P3_training" / n, simple=True) for n in
             ("maze_3d_tunnels.txt", "maze_7x7_seed22.txt", "kitchen2.txt")] + \
            [load_room_file(ROOMS / "P2_training" / "tightcorridor.txt", simple=True)]
    n, steps, seed = 96, 500, 13
    orooms = [oracle.OracleRoom(r.grid, 2) for r in rooms]
    ov = oracle.OracleSimpleVec(n, orooms, L, 0.25, seed, 7, True)
    g = GpuSimple(n, rooms, L, seed, 7, True, lanes)
    assert np.array_equal(g.reset().view(np.uint32), ov.reset().view(np.uint32))
    rng = np.random.default_rng(1)
    n_done = 0
    for t in range(steps):
        a = rng.integers(0, 6, size=n)
        ov.step(a)
        g.step(a)
        assert np.array_equal(g.term, ov.terminated) and np.array_equal(g.trunc, ov.truncated), t
        assert np.array_equal(g.obs.view(np.uint32), ov.obs.view(np.uint32)), t
        assert np.array_equal(g.reward64, ov.reward), t
        done = (ov.terminated | ov.truncated).astype(bool)
        if done.any():
            assert np.array_equal(g.tobs[done].view(np.uint32), ov.terminal_obs[done].view(np.uint32)), t
            n_done += int(done.sum())
        if t % 25 == 0:
            s, so = g.state(), ov.state()
            assert np.array_equal(s[:, [0, 1, 2, 3, 4, 5, 6, 10, 13, 14]].astype(np.int64), so), t
    assert n_done > 20


def test_simple_scalar_facade(simple_traces, capsys):
    from pathlib import Path

    from envs.simpleEnv import GridAgent
    c = simple_traces[0]
    p = ROOMS / c["room"]
    env = GridAgent(room_path=str(p.parent), local_map_length=c["L"])
    env.rooms = [Path(p)]
    random.seed(c["seed"])                      # the trace was generated with random.seed(seed); reset()
    obs, _ = env.reset()
    assert (env.x, env.y, env.z) == tuple(c["start"]) and (env.gx, env.gy, env.gz) == tuple(c["goal"])
    assert np.array_equal(obs.view(np.uint32), c["obs"][0].view(np.uint32)) and obs.shape == (6 * c["L"] + 7,)
    for t in range(200):
        o, r, term, trunc, _ = env.step(int(c["actions"][t]))
        assert r == c["reward"][t] and np.array_equal(o.view(np.uint32), c["obs"][t + 1].view(np.uint32))
    env.close()
    capsys.readouterr()
