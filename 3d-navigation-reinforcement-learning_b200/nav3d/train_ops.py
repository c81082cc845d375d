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
                       entropy: Optional[torch.Tensor] = None,
                       step_offset: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """logits f32 [N, A] -> (actions int64 [N], log_prob f32 [N]); ``step`` (+ the int32 device scalar ``step_offset``,
        read by the kernel) indexes the env's Philox policy stream."""
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
            if step_offset is not None and (step_offset.dtype != torch.int32 or step_offset.numel() != 1 or not step_offset.is_cuda):
                raise ValueError("step_offset must be a CUDA int32 tensor with one element")
            check(self._lib.nav3d_sample_actions(_ptr(logits), n, a, self.seed, self.env_id0, int(step) & 0xFFFFFFFF,
                                                 _ptr(step_offset), int(bool(greedy)), _ptr(actions), _ptr(log_prob),
                                                 _ptr(entropy), _stream(dev)))
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


def lstm_prepare_stream(stream: "torch.cuda.Stream") -> None:
    """Create the fused LSTM's cuBLAS handle + workspace for `stream` (include/nav3d.h nav3d_lstm_prepare): to be called
    before a CUDA-graph capture on that stream, after which fused_lstm is capturable on it."""
    with torch.cuda.device(stream.device):
        check(_lib.load().nav3d_lstm_prepare(C.c_void_p(stream.cuda_stream)))


class _FusedLSTMFn(torch.autograd.Function):
    """``nav3d_lstm_forward`` / ``nav3d_lstm_backward`` as one autograd node: (x, w_ih, w_hh, b_ih, b_hh, h0, c0, starts)
    -> (h_all, h_last, c_last).  Gradients flow to the four parameters only (see include/nav3d.h)."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, h0, c0, starts, tf32):
        lib = _lib.load()
        S, B, F = x.shape
        H = w_hh.shape[1]
        dev = x.device
        x, h0, c0, starts = x.contiguous(), h0.contiguous(), c0.contiguous(), starts.contiguous()
        w_ih_c, w_hh_c, b_ih_c, b_hh_c = w_ih.contiguous(), w_hh.contiguous(), b_ih.contiguous(), b_hh.contiguous()
        gates = torch.empty((S, B, 4 * H), dtype=torch.float32, device=dev)
        h_in = torch.empty((S, B, H), dtype=torch.float32, device=dev)
        h_all = torch.empty((S, B, H), dtype=torch.float32, device=dev)
        c_all = torch.empty((S, B, H), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.nav3d_lstm_forward(_ptr(x), _ptr(w_ih_c), _ptr(w_hh_c), _ptr(b_ih_c), _ptr(b_hh_c), _ptr(h0), _ptr(c0),
                                         _ptr(starts), S, B, F, H, int(tf32), _ptr(gates), _ptr(h_in), _ptr(h_all),
                                         _ptr(c_all), _stream(dev)))
        ctx.save_for_backward(x, w_hh_c, c0, starts, h_in, c_all, gates)
        ctx.tf32 = int(tf32)
        ctx.set_materialize_grads(False)
        h_last, c_last = h_all[-1].clone(), c_all[-1].clone()
        ctx.mark_non_differentiable(h_last, c_last)
        return h_all, h_last, c_last

    @staticmethod
    def backward(ctx, dh_all, _dh_last, _dc_last):
        if dh_all is None:
            return (None,) * 9
        lib = _lib.load()
        x, w_hh, c0, starts, h_in, c_all, gates = ctx.saved_tensors
        S, B, F = x.shape
        H = w_hh.shape[1]
        dev = x.device
        dh_all = dh_all.contiguous()
        dw_ih = torch.empty((4 * H, F), dtype=torch.float32, device=dev)
        dw_hh = torch.empty((4 * H, H), dtype=torch.float32, device=dev)
        db = torch.empty(4 * H, dtype=torch.float32, device=dev)
        scratch = torch.empty(2 * B * H + S * B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            # `gates` is consumed (overwritten with the pre-activation gradients): this node can be differentiated once
            check(lib.nav3d_lstm_backward(_ptr(x), _ptr(w_hh), _ptr(c0), _ptr(starts), _ptr(h_in), _ptr(c_all), _ptr(gates),
                                          _ptr(dh_all), S, B, F, H, ctx.tf32, _ptr(dw_ih), _ptr(dw_hh), _ptr(db),
                                          _ptr(scratch), _stream(dev)))
        return None, dw_ih, dw_hh, db, db, None, None, None, None


def fused_lstm(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, h0, c0, starts: torch.Tensor, tf32: Optional[bool] = None):
    """x f32 [S,B,F], h0/c0 f32 [B,H], starts u8 [S,B] -> (h_all [S,B,H], h_last [B,H], c_last [B,H])."""
    _need_cuda(x, w_ih, w_hh, h0, c0, starts)
    if x.dtype != torch.float32 or starts.dtype != torch.uint8:
        raise ValueError("fused_lstm: x must be float32 and starts uint8")
    if tf32 is None:
        tf32 = torch.backends.cuda.matmul.allow_tf32
    return _FusedLSTMFn.apply(x, w_ih, w_hh, b_ih, b_hh, h0, c0, starts, bool(tf32))
