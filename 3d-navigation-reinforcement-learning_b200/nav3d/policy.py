"""``RecurrentActorCritic`` — the network the reference trains: sb3-contrib's ``MlpLstmPolicy`` as configured by
``train/Grid_Train.py:68-88`` and built at ``:198-205`` (``policy_kwargs = dict(net_arch=dict(pi=[256,256,128],
vf=[256,256,128]), lstm_hidden_size=256, n_lstm_layers=1)``).

sb3-contrib is a third-party dependency the reference does not vendor (and it is not installed here), so this module
restates the published architecture of ``RecurrentActorCriticPolicy`` rather than following reference source lines:

* features = the flat 80-float observation (``Flatten`` extractor);
* ``lstm_actor``: ``nn.LSTM(features, H, n_layers)``; critic memory is one of: its own ``lstm_critic`` (default:
  ``shared_lstm=False, enable_critic_lstm=True``), the actor's LSTM (``shared_lstm=True``; the actor's output is detached
  for the value head), or no memory (``critic = nn.Linear(features, H)``);
* ``mlp_extractor.policy_net`` / ``.value_net``: ``Linear -> Tanh`` stacks given by ``net_arch``;
* ``action_net = Linear(last_pi, 6)`` (categorical logits), ``value_net = Linear(last_vf, 1)``;
* orthogonal init with gains sqrt(2) (MLPs), 0.01 (``action_net``), 1 (``value_net``); LSTMs keep PyTorch's default init;
* hidden/cell states are zeroed where ``episode_starts`` is set, *before* the step that consumes them.

Parameter names match sb3-contrib's ``state_dict`` keys, so a checkpoint's ``policy.pth`` has the layout SB3 users expect.

The MLP GEMMs run in torch (cuBLAS on the tensor cores); the env step is the hand-written CUDA path.  On a CUDA device a
sequence goes through the library's fused LSTM (``nav3d_lstm_forward/backward``: cuBLAS GEMMs + one hand-written kernel per
timestep, the episode-start mask as an operand), with the critic branch on a second stream; single steps, CPU tensors and
multi-layer LSTMs use ``torch.nn.LSTM`` over every stretch of timesteps that contains no episode start — either way the
per-step Python loop SB3 falls back to whenever a rollout contains a reset is avoided."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

LSTMState = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]   # (h_pi, c_pi, h_vf, c_vf), each [layers, B, H]


def _mlp(in_dim: int, sizes: Sequence[int]) -> nn.Sequential:
    layers: List[nn.Module] = []
    d = in_dim
    for s in sizes:
        layers += [nn.Linear(d, s), nn.Tanh()]
        d = s
    return nn.Sequential(*layers)


class _MlpExtractor(nn.Module):
    def __init__(self, in_dim: int, pi: Sequence[int], vf: Sequence[int]):
        super().__init__()
        self.policy_net = _mlp(in_dim, pi)
        self.value_net = _mlp(in_dim, vf)
        self.latent_dim_pi = pi[-1] if len(pi) else in_dim
        self.latent_dim_vf = vf[-1] if len(vf) else in_dim


class RecurrentActorCritic(nn.Module):
    def __init__(self, obs_dim: int = 80, n_actions: int = 6, net_arch: Optional[Dict[str, Sequence[int]]] = None,
                 lstm_hidden_size: int = 256, n_lstm_layers: int = 1, shared_lstm: bool = False,
                 enable_critic_lstm: bool = True, ortho_init: bool = True):
        super().__init__()
        if shared_lstm and enable_critic_lstm:
            # sb3-contrib asserts the same: a shared LSTM and a critic LSTM are mutually exclusive.  The reference never
            # passes either flag, so it gets the default (separate critic LSTM).
            enable_critic_lstm = False
        net_arch = dict(net_arch or dict(pi=[64, 64], vf=[64, 64]))
        self.obs_dim, self.n_actions = int(obs_dim), int(n_actions)
        self.net_arch = {k: list(v) for k, v in net_arch.items()}
        self.lstm_hidden_size, self.n_lstm_layers = int(lstm_hidden_size), int(n_lstm_layers)
        self.shared_lstm, self.enable_critic_lstm = bool(shared_lstm), bool(enable_critic_lstm)
        H = self.lstm_hidden_size
        self.lstm_actor = nn.LSTM(self.obs_dim, H, num_layers=self.n_lstm_layers)
        self.lstm_critic = nn.LSTM(self.obs_dim, H, num_layers=self.n_lstm_layers) if self.enable_critic_lstm else None
        self.critic = nn.Linear(self.obs_dim, H) if not (self.shared_lstm or self.enable_critic_lstm) else None
        self.mlp_extractor = _MlpExtractor(H, self.net_arch.get("pi", []), self.net_arch.get("vf", []))
        self.action_net = nn.Linear(self.mlp_extractor.latent_dim_pi, self.n_actions)
        self.value_net = nn.Linear(self.mlp_extractor.latent_dim_vf, 1)
        self.two_streams = True
        self.fused_lstm = True
        self._streams: Dict[str, "torch.cuda.Stream"] = {}
        if ortho_init:
            for module, gain in ((self.mlp_extractor, math.sqrt(2.0)), (self.action_net, 0.01), (self.value_net, 1.0)):
                for m in module.modules():
                    if isinstance(m, nn.Linear):
                        nn.init.orthogonal_(m.weight, gain=gain)
                        nn.init.zeros_(m.bias)

    # ---- constructor kwargs, for checkpoints -------------------------------------------------------------------
    def kwargs(self) -> dict:
        return dict(obs_dim=self.obs_dim, n_actions=self.n_actions, net_arch=self.net_arch,
                    lstm_hidden_size=self.lstm_hidden_size, n_lstm_layers=self.n_lstm_layers,
                    shared_lstm=self.shared_lstm, enable_critic_lstm=self.enable_critic_lstm)

    # ---- state -------------------------------------------------------------------------------------------------
    def initial_state(self, batch: int, device=None, dtype=torch.float32) -> LSTMState:
        device = device if device is not None else self.action_net.weight.device
        z = lambda: torch.zeros((self.n_lstm_layers, batch, self.lstm_hidden_size), device=device, dtype=dtype)  # noqa: E731
        return z(), z(), z(), z()

    @staticmethod
    def _masked(state: Tuple[torch.Tensor, torch.Tensor], starts: torch.Tensor):
        keep = (1.0 - starts.to(state[0].dtype)).view(1, -1, 1)
        return state[0] * keep, state[1] * keep

    def _run_lstm(self, lstm: nn.LSTM, x: torch.Tensor, state, starts: torch.Tensor, cuts: Sequence[int]):
        """x [S,B,F], starts [S,B] (1 = the state entering step t is zeroed), cuts = sorted timesteps at which a mask
        has to be applied (0 is always one).

        CUDA tensors go through the library's fused LSTM (``nav3d_lstm_forward/backward``: the episode-start mask is an
        operand, so nothing is cut, and the backward avoids cuDNN's bulk gate-gradient pass — DESIGN.md §6b), single rollout
        steps included; it is capturable in a CUDA graph once ``train_ops.lstm_prepare_stream`` has run for the capture
        stream.  CPU tensors and multi-layer LSTMs use torch's LSTM, run over every stretch between cuts."""
        if self.fused_lstm and x.is_cuda and lstm.num_layers == 1 and x.dtype == torch.float32:
            from .train_ops import fused_lstm
            y, h, c = fused_lstm(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, state[0][0],
                                 state[1][0], starts)
            return y, (h.unsqueeze(0), c.unsqueeze(0))
        outs = []
        S = x.shape[0]
        cuts = list(cuts)
        for i, t0 in enumerate(cuts):
            t1 = cuts[i + 1] if i + 1 < len(cuts) else S
            state = RecurrentActorCritic._masked(state, starts[t0])
            y, state = lstm(x[t0:t1], state)
            outs.append(y)
        return (outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)), state

    # ---- forward passes ----------------------------------------------------------------------------------------
    def _critic_branch(self, obs, h_vf, c_vf, starts, cuts):
        lat_vf, (h_vf, c_vf) = self._run_lstm(self.lstm_critic, obs, (h_vf, c_vf), starts, cuts)
        return self.value_net(self.mlp_extractor.value_net(lat_vf)).squeeze(-1), h_vf, c_vf

    def forward_sequence(self, obs: torch.Tensor, state: LSTMState, starts: torch.Tensor,
                         cuts: Optional[Sequence[int]] = None):
        """obs [S,B,F], starts [S,B] -> logits [S,B,A], values [S,B], final state.  ``cuts=None`` = mask at every step
        (always correct); pass the timesteps that can hold an episode start to let cuDNN run whole stretches.

        With a separate critic LSTM on a CUDA device the critic branch (LSTM + value MLP) runs on a second stream: the two
        recurrences are independent chains of small per-timestep kernels, so running them side by side (forward, and —
        because autograd replays each op on its forward stream — backward too) hides one chain behind the other."""
        if cuts is None:
            cuts = range(obs.shape[0])
        h_pi, c_pi, h_vf, c_vf = state
        side = None
        if self.lstm_critic is not None and obs.is_cuda and self.two_streams and not torch.cuda.is_current_stream_capturing():
            main = torch.cuda.current_stream(obs.device)
            side = self._side_stream(obs.device)
            # the critic's parameters accumulate their gradients on the side stream by design
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                values, h_vf, c_vf = self._critic_branch(obs, h_vf, c_vf, starts, cuts)
        elif self.lstm_critic is not None:
            values, h_vf, c_vf = self._critic_branch(obs, h_vf, c_vf, starts, cuts)
        lat_pi, (h_pi, c_pi) = self._run_lstm(self.lstm_actor, obs, (h_pi, c_pi), starts, cuts)
        logits = self.action_net(self.mlp_extractor.policy_net(lat_pi))
        if side is not None:
            main.wait_stream(side)
            for t in (values, h_vf, c_vf):
                t.record_stream(main)               # allocated on the side stream, consumed on the main one
        elif self.lstm_critic is None:
            if self.shared_lstm:
                lat_vf, h_vf, c_vf = lat_pi.detach(), h_pi.detach(), c_pi.detach()
            else:
                lat_vf = self.critic(obs)
            values = self.value_net(self.mlp_extractor.value_net(lat_vf)).squeeze(-1)
        return logits, values, (h_pi, c_pi, h_vf, c_vf)

    def _side_stream(self, device):
        key = str(device)
        if key not in self._streams:
            self._streams[key] = torch.cuda.Stream(device=device)
        return self._streams[key]

    def forward_step(self, obs: torch.Tensor, state: LSTMState, starts: torch.Tensor):
        """One timestep: obs [B,F], starts [B] -> logits [B,A], values [B], new state."""
        logits, values, state = self.forward_sequence(obs.unsqueeze(0), state, starts.unsqueeze(0), (0,))
        return logits[0], values[0], state

    def values_step(self, obs: torch.Tensor, state: LSTMState, starts: torch.Tensor) -> torch.Tensor:
        """Critic only (``predict_values``): used for the last observation of a rollout and for time-limit bootstraps."""
        obs, starts = obs.unsqueeze(0), starts.unsqueeze(0)
        h_pi, c_pi, h_vf, c_vf = state
        if self.lstm_critic is not None:
            lat_vf, _ = self._run_lstm(self.lstm_critic, obs, (h_vf, c_vf), starts, (0,))
        elif self.shared_lstm:
            lat_vf, _ = self._run_lstm(self.lstm_actor, obs, (h_pi, c_pi), starts, (0,))
        else:
            lat_vf = self.critic(obs)
        return self.value_net(self.mlp_extractor.value_net(lat_vf)).squeeze(-1)[0]
