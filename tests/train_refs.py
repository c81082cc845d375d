"""Plain-PyTorch fp32 restatements of the rollout-loop kernels (``csrc/nav3d_train.cu``) and an oracle-backed stand-in
for ``nav3d.BatchedCubicEnv``.  TEST INFRASTRUCTURE ONLY: the GPU tests check the CUDA kernels against these, and the CPU
tests run the trainer's host logic on them (the product has no CPU path)."""
from types import SimpleNamespace

import numpy as np
import torch


def gae_reference(rewards, values, starts, last_values, last_dones, gamma, lam):
    """SB3's compute_returns_and_advantage, time-major, fp32."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    gae = torch.zeros_like(last_values)
    for t in reversed(range(T)):
        if t == T - 1:
            nnt, nv = 1.0 - last_dones.float(), last_values
        else:
            nnt, nv = 1.0 - starts[t + 1].float(), values[t + 1]
        delta = rewards[t] + gamma * nv * nnt - values[t]
        gae = delta + gamma * lam * nnt * gae
        adv[t] = gae
    return adv, adv + values


class TorchOps:
    """Same interface as nav3d.train_ops.DeviceOps, in torch, on any device."""

    def __init__(self, seed=0):
        self.gen = torch.Generator().manual_seed(seed)

    def sample_actions(self, logits, step, greedy=False, actions=None, log_prob=None, entropy=None, step_offset=None):
        lp = torch.log_softmax(logits, dim=-1)
        a = lp.argmax(-1) if greedy else torch.multinomial(lp.exp().cpu(), 1, generator=self.gen).squeeze(-1).to(logits.device)
        l = lp.gather(-1, a.unsqueeze(-1)).squeeze(-1)
        if actions is not None:
            actions.copy_(a)
        if log_prob is not None:
            log_prob.copy_(l)
        if entropy is not None:
            entropy.copy_(-(lp.exp() * lp).sum(-1))
        return (actions if actions is not None else a), (log_prob if log_prob is not None else l)

    def gae(self, rewards, values, episode_starts, last_values, last_dones, gamma, gae_lambda, advantages, returns):
        a, r = gae_reference(rewards, values, episode_starts, last_values, last_dones, gamma, gae_lambda)
        advantages.copy_(a)
        returns.copy_(r)


class OracleBatchedEnv:
    """``nav3d.BatchedCubicEnv``'s interface on the CPU oracle (``oracle.c_oracle.OracleVec``), torch CPU tensors."""

    def __init__(self, rooms, num_envs, local_map_length=10, seed=0, crash_penalty=-2.0):
        from oracle import c_oracle
        self.ov = c_oracle.OracleVec(num_envs, [c_oracle.OracleRoom(r.grid, -2) for r in rooms], local_map_length,
                                     crash_penalty, seed, 0, True)
        self.num_envs = num_envs
        self.device = torch.device("cpu")
        self.engine = SimpleNamespace(env_id0=0)
        self._obs = torch.zeros((num_envs, 80), dtype=torch.float32)
        self._eps = torch.zeros((num_envs, 8), dtype=torch.int32)

    def reset(self, picks=None):
        self._obs.copy_(torch.from_numpy(self.ov.reset(picks)))
        return self._obs

    def step(self, actions, out_obs=None):
        ov = self.ov
        ov.step(torch.as_tensor(actions).cpu().numpy())
        obs = self._obs if out_obs is None else out_obs
        obs.copy_(torch.from_numpy(ov.obs))
        term, trunc = torch.from_numpy(ov.terminated.copy()), torch.from_numpy(ov.truncated.copy())
        done = (ov.terminated | ov.truncated).astype(bool)
        eps = self._eps.numpy()
        eps[done, 0] = ov.ep_ret[done].astype(np.float32).view(np.int32)
        eps[done, 1], eps[done, 2], eps[done, 3] = ov.ep_len[done], ov.ep_bumps[done], ov.ep_visited[done]
        eps[done, 6], eps[done, 7] = ov.terminated[done], ov.truncated[done]
        info = SimpleNamespace(terminated=term, truncated=trunc, terminal_observation=torch.from_numpy(ov.terminal_obs.copy()),
                               episodes=self._eps)
        return obs, torch.from_numpy(ov.reward.astype(np.float32)), torch.from_numpy(done), info

    def close(self):
        pass
