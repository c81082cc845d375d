"""Policy evaluation on a batched env and the periodic-evaluation callback of the reference's training loop.

Replaces stable-baselines3's ``evaluate_policy`` / ``EvalCallback`` as the reference uses them
(``train/Grid_Train.py:218-226``: ``eval_freq = EVAL_FREQ // NUM_ENVS`` vec-env steps, ``n_eval_episodes=10``,
``deterministic=True``, best model saved under ``best_model_save_path``, results logged under ``log_path``;
same at ``train/Train_Further.py:162-170``).  SB3 is third-party and un-vendored; the behaviour restated here is its
documented one: episodes are shared out over the eval envs as evenly as possible (env i plays
``(n_eval_episodes + i) // n_envs`` episodes), LSTM states are reset at episode starts, the best mean reward so far
triggers ``best_model.zip``, and every evaluation is appended to ``evaluations.npz`` (timesteps, results, ep_lengths)."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch


@torch.no_grad()
def evaluate_policy(model, env, n_eval_episodes: int = 10, deterministic: bool = True,
                    return_episode_stats: bool = False, max_steps: Optional[int] = None):
    """Play ``n_eval_episodes`` episodes of ``model`` on the auto-resetting batched ``env``.

    Returns ``(mean_reward, std_reward)``, or with ``return_episode_stats`` a dict of per-episode arrays
    (``r``, ``l``, ``bumps``, ``visited``, ``total_free``, ``terminated``)."""
    n = env.num_envs
    targets = np.array([(n_eval_episodes + i) // n for i in range(n)], dtype=np.int64)
    counts = np.zeros(n, dtype=np.int64)
    dev = model.device
    obs = env.reset()
    state = model.policy.initial_state(n, dev)
    starts = torch.ones(n, dtype=torch.uint8, device=dev)
    rows: List[np.ndarray] = []
    steps = 0
    while (counts < targets).any():
        actions, state = model.predict(obs, state=state, episode_start=starts, deterministic=deterministic)
        obs, _, dones, info = env.step(actions)
        starts = dones.to(torch.uint8)
        steps += 1
        d = dones.cpu().numpy()
        if d.any():
            eps = info.episodes.cpu().numpy()
            for i in np.nonzero(d)[0]:
                if counts[i] < targets[i]:
                    counts[i] += 1
                    rows.append(eps[i].copy())
        if max_steps is not None and steps >= max_steps:
            break
    ep = np.stack(rows) if rows else np.zeros((0, 8), dtype=np.int32)
    r = ep[:, 0].copy().view(np.float32).astype(np.float64)
    if return_episode_stats:
        return dict(r=r, l=ep[:, 1].astype(np.int64), bumps=ep[:, 2].astype(np.int64), visited=ep[:, 3].astype(np.int64),
                    total_free=ep[:, 4].astype(np.int64), room=ep[:, 5].astype(np.int64), terminated=ep[:, 6].astype(bool),
                    truncated=ep[:, 7].astype(bool))
    return (float(r.mean()) if r.size else float("nan")), (float(r.std()) if r.size else float("nan"))


class EvalCallback:
    """Evaluate every ``eval_freq`` calls of ``on_step`` (one call per vec-env step, as in SB3) and keep the best model."""

    def __init__(self, eval_env, best_model_save_path=None, log_path=None, eval_freq: int = 10000,
                 n_eval_episodes: int = 5, deterministic: bool = True, render: bool = False, verbose: int = 1):
        self.eval_env = eval_env
        self.best_model_save_path = Path(best_model_save_path) if best_model_save_path is not None else None
        self.log_path = Path(log_path) / "evaluations" if log_path is not None else None
        self.eval_freq, self.n_eval_episodes, self.deterministic = int(eval_freq), int(n_eval_episodes), bool(deterministic)
        self.verbose = verbose
        self.n_calls = 0
        self.best_mean_reward = -np.inf
        self.last_mean_reward = -np.inf
        self.evaluations_timesteps: List[int] = []
        self.evaluations_results: List[List[float]] = []
        self.evaluations_length: List[List[int]] = []

    def init_callback(self, model) -> None:
        if self.best_model_save_path is not None:
            self.best_model_save_path.mkdir(parents=True, exist_ok=True)
        if self.log_path is not None:
            self.log_path.parent.mkdir(parents=True, exist_ok=True)

    def on_step(self, model) -> bool:
        self.n_calls += 1
        if self.eval_freq > 0 and self.n_calls % self.eval_freq == 0:
            self.evaluate(model)
        return True

    def evaluate(self, model) -> Tuple[float, float]:
        st = evaluate_policy(model, self.eval_env, self.n_eval_episodes, self.deterministic, return_episode_stats=True)
        mean_r, std_r = float(st["r"].mean()), float(st["r"].std())
        self.last_mean_reward = mean_r
        if getattr(model, "rank", 0) == 0:
            if self.log_path is not None:
                self.evaluations_timesteps.append(int(model.num_timesteps))
                self.evaluations_results.append([float(v) for v in st["r"]])
                self.evaluations_length.append([int(v) for v in st["l"]])
                np.savez(self.log_path, timesteps=np.array(self.evaluations_timesteps),
                         results=np.array(self.evaluations_results), ep_lengths=np.array(self.evaluations_length))
            if self.verbose:
                print(f"Eval num_timesteps={model.num_timesteps}, episode_reward={mean_r:.2f} +/- {std_r:.2f}")
                print(f"Episode length: {st['l'].mean():.2f} +/- {st['l'].std():.2f}")
        if mean_r > self.best_mean_reward:
            if self.verbose and getattr(model, "rank", 0) == 0:
                print("New best mean reward!")
            if self.best_model_save_path is not None:
                model.save(self.best_model_save_path / "best_model")
            self.best_mean_reward = mean_r
        return mean_r, std_r
