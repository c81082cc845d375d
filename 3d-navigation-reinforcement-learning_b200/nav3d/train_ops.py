"""Python face of the rollout-loop kernels in ``csrc/nav3d_train.cu`` (C ABI: ``nav3d_sample_actions``, ``nav3d_gae``).

CUDA tensors only — like the env step, these have no CPU implementation in the product (the torch restatements that the
tests check them against live in ``tests/``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nav3d.train_ops needs CUDA tensors (B200, sm_100a); there is no CPU fallback")


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class DeviceOps:
    """The two per-env kernels bound to one Philox key; ``RecurrentPPO`` takes an object with this interface."""

    def __init__(self, seed: int = 0, env_id0: int = 0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.env_id0 = int(env_id0) & 0xFFFFFFFF
        self._lib = _lib.load()

    def sample_actions(self, logits: torch.Tensor, step: int, greedy: bool = False,
                       actions: Optional[torch.Tensor] = None, log_prob: Optional[torch.Tensor] = None,
                       entropy: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """logits f32 [N, A] -> (actions int64 [N], log_prob f32 [N]); ``step`` indexes the env's Philox policy stream."""
        _need_cuda(logits)
        if logits.dtype != torch.float32 or logits.dim() != 2:
            raise ValueError("logits must be a float32 [N, A] tensor")
        logits = logits.contiguous()
        n, a = logits.shape
        dev = logits.device
        if actions is None:
            actions = torch.empty(n, dtype=torch.int64, device=dev)
        if log_prob is None:
            log_prob = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.nav3d_sample_actions(_ptr(logits), n, a, self.seed, self.env_id0, int(step) & 0xFFFFFFFF,
                                                 int(bool(greedy)), _ptr(actions), _ptr(log_prob), _ptr(entropy),
                                                 _stream(dev)))
        return actions, log_prob

    def gae(self, rewards: torch.Tensor, values: torch.Tensor, episode_starts: torch.Tensor, last_values: torch.Tensor,
            last_dones: torch.Tensor, gamma: float, gae_lambda: float, advantages: torch.Tensor,
            returns: torch.Tensor) -> None:
        """Time-major [T, N] rollout -> ``advantages``, ``returns`` (both written in place)."""
        _need_cuda(rewards, values, episode_starts, last_values, last_dones, advantages, returns)
        T, N = rewards.shape
        for t, dt, shp in ((rewards, torch.float32, (T, N)), (values, torch.float32, (T, N)),
                           (episode_starts, torch.uint8, (T, N)), (last_values, torch.float32, (N,)),
                           (last_dones, torch.uint8, (N,)), (advantages, torch.float32, (T, N)),
                           (returns, torch.float32, (T, N))):
            if t.dtype != dt or tuple(t.shape) != shp or not t.is_contiguous():
                raise ValueError(f"gae: expected a contiguous {dt} tensor of shape {shp}, got {t.dtype} {tuple(t.shape)}")
        dev = rewards.device
        with torch.cuda.device(dev):
            check(self._lib.nav3d_gae(_ptr(rewards), _ptr(values), _ptr(episode_starts), _ptr(last_values),
                                      _ptr(last_dones), float(gamma), float(gae_lambda), T, N, _ptr(advantages),
                                      _ptr(returns), _stream(dev)))
