"""Adapter used by the -m gpu tests: drives the product (nav3d.Engine -> C ABI -> CUDA kernels) and exposes numpy views
with the same attribute names the lock-step helpers expect."""
import numpy as np
import torch

from nav3d import Engine


class GpuEngine:
    def __init__(self, n, rooms, L=4, crash_penalty=-2.0, seed=0, env_id0=0, auto_reset=True, lanes=0, device=0):
        self.n = n
        self.eng = Engine(n, rooms, local_map_length=L, crash_penalty=crash_penalty, auto_reset=auto_reset, seed=seed,
                          env_id0=env_id0, device=device, lanes_per_env=lanes)
        dev = self.eng.device
        self.d_obs = torch.full((n, 80), float("nan"), dtype=torch.float32, device=dev)
        self.d_reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.d_reward64 = torch.zeros(n, dtype=torch.float64, device=dev)
        self.d_term = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.d_trunc = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.d_tobs = torch.zeros((n, 80), dtype=torch.float32, device=dev)
        self.d_eps = torch.zeros((n, 8), dtype=torch.int32, device=dev)

    def n_free(self, r):
        return self.eng.room_free[r]

    def _pull(self):
        self.obs = self.d_obs.cpu().numpy()
        self.reward = self.d_reward.cpu().numpy()
        self.reward64 = self.d_reward64.cpu().numpy()
        self.term = self.d_term.cpu().numpy()
        self.trunc = self.d_trunc.cpu().numpy()
        self.tobs = self.d_tobs.cpu().numpy()
        self.eps = self.d_eps.cpu().numpy()

    def reset(self, picks=None, env_ids=None):
        p = None if picks is None else torch.as_tensor(np.asarray(picks, dtype=np.int32))
        ids = None if env_ids is None else torch.as_tensor(np.asarray(env_ids, dtype=np.int32))
        self.eng.reset(self.d_obs, env_ids=ids, picks=p)
        self._pull()
        return self.obs

    def step(self, actions, pull=True):
        a = torch.as_tensor(np.asarray(actions, dtype=np.int64)).to(self.eng.device)
        self.eng.step(a, self.d_obs, self.d_reward, self.d_term, self.d_trunc, reward64=self.d_reward64,
                      terminal_obs=self.d_tobs, episodes=self.d_eps)
        if pull:
            self._pull()

    def state(self):
        return self.eng.get_state().cpu().numpy()

    def grid(self, env):
        return self.eng.get_grid(env)
