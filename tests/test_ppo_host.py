"""Host logic of the LSTM-PPO trainer on CPU: the network against a step-by-step evaluation of itself,
the rollout/update bookkeeping on an oracle-backed env with torch reference ops, checkpoints, evaluation."""
import math

import numpy as np
import pytest
import torch

from nav3d.evaluation import EvalCallback, evaluate_policy
from nav3d.policy import RecurrentActorCritic
from nav3d.ppo import RecurrentPPO
from nav3d.rooms import rooms_from_grids

from train_refs import OracleBatchedEnv, TorchOps, gae_reference

REF_POLICY_KWARGS = dict(net_arch=dict(pi=[256, 256, 128], vf=[256, 256, 128]), lstm_hidden_size=256, n_lstm_layers=1)


def tiny_rooms():
    """A 6x5x5 hollow box (36 free cells: episodes truncate after 36 steps) and a 5x5x4 one with a pillar."""
    g1 = np.zeros((6, 5, 5), dtype=np.int8)
    g1[0], g1[-1], g1[:, 0], g1[:, -1], g1[:, :, 0], g1[:, :, -1] = -2, -2, -2, -2, -2, -2
    g2 = np.zeros((5, 5, 4), dtype=np.int8)
    g2[0], g2[-1], g2[:, 0], g2[:, -1], g2[:, :, 0], g2[:, :, -1] = -2, -2, -2, -2, -2, -2
    g2[2, 2, 1] = -2
    return rooms_from_grids([g1, g2])


def test_policy_layout_matches_sb3_names():
    pol = RecurrentActorCritic(**REF_POLICY_KWARGS)
    keys = set(pol.state_dict().keys())
    for k in ("lstm_actor.weight_ih_l0", "lstm_actor.weight_hh_l0", "lstm_critic.weight_ih_l0", "lstm_critic.bias_hh_l0",
              "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.4.bias", "mlp_extractor.value_net.2.weight",
              "action_net.weight", "value_net.bias"):
        assert k in keys, k
    assert pol.lstm_actor.weight_ih_l0.shape == (1024, 80) and pol.action_net.weight.shape == (6, 128)
    assert pol.value_net.weight.shape == (1, 128) and pol.mlp_extractor.policy_net[2].weight.shape == (256, 256)
    n_params = sum(p.numel() for p in pol.parameters())
    assert n_params == 2 * (4 * 256 * (80 + 256) + 8 * 256) + 2 * (256 * 256 + 256 + 256 * 256 + 256 + 256 * 128 + 128) + 6 * 128 + 6 + 128 + 1
    # orthogonal init: rows of action_net are orthogonal with norm 0.01, value_net has norm 1
    w = pol.action_net.weight.detach()
    assert torch.allclose(w @ w.T, 1e-4 * torch.eye(6), atol=1e-7)
    assert abs(float(pol.value_net.weight.norm()) - 1.0) < 1e-5


@pytest.mark.parametrize("variant", ["separate", "shared", "no_critic_lstm"])
def test_sequence_forward_equals_stepwise(variant):
    torch.manual_seed(3)
    kw = dict(net_arch=dict(pi=[32, 16], vf=[24]), lstm_hidden_size=20, n_lstm_layers=2)
    if variant == "shared":
        kw.update(shared_lstm=True, enable_critic_lstm=False)
    elif variant == "no_critic_lstm":
        kw.update(shared_lstm=False, enable_critic_lstm=False)
    pol = RecurrentActorCritic(obs_dim=11, n_actions=5, **kw).double()
    S, B = 9, 7
    obs = torch.randn(S, B, 11, dtype=torch.float64)
    starts = (torch.rand(S, B) < 0.2).to(torch.uint8)
    starts[3] = 0
    starts[4] = 0
    state0 = tuple(torch.randn(2, B, 20, dtype=torch.float64) for _ in range(4))
    # reference: one step at a time
    st = state0
    ref_l, ref_v = [], []
    for t in range(S):
        l, v, st = pol.forward_step(obs[t], st, starts[t])
        ref_l.append(l); ref_v.append(v)
    ref_l, ref_v = torch.stack(ref_l), torch.stack(ref_v)
    cuts = [0] + [t for t in range(1, S) if bool(starts[t].any())]
    assert 3 not in cuts and 4 not in cuts
    for c in (None, cuts):
        l, v, st2 = pol.forward_sequence(obs, state0, starts, c)
        assert torch.allclose(l, ref_l, atol=1e-12) and torch.allclose(v, ref_v, atol=1e-12)
        for a, b in zip(st, st2):
            assert torch.allclose(a, b, atol=1e-12)
    # an episode start really cuts the dependence on the past
    starts2 = starts.clone(); starts2[5] = 1
    other = tuple(s + 1.0 for s in state0)
    la, _, _ = pol.forward_sequence(obs, state0, starts2)
    lb, _, _ = pol.forward_sequence(obs, other, starts2)
    assert torch.allclose(la[5:], lb[5:], atol=1e-12) and not torch.allclose(la[:5], lb[:5], atol=1e-6)
    # values_step is the critic half of forward_step
    _, v1, _ = pol.forward_step(obs[0], state0, starts[0])
    assert torch.allclose(pol.values_step(obs[0], state0, starts[0]), v1, atol=1e-12)


def test_gae_reference_against_scalar_recursion():
    rng = np.random.default_rng(0)
    T, N, g, lam = 13, 5, 0.99, 0.95
    r, v = rng.normal(size=(T, N)).astype(np.float32), rng.normal(size=(T, N)).astype(np.float32)
    s = (rng.random((T, N)) < 0.2).astype(np.uint8)
    lv, ld = rng.normal(size=N).astype(np.float32), (rng.random(N) < 0.5).astype(np.uint8)
    adv, ret = gae_reference(torch.from_numpy(r), torch.from_numpy(v), torch.from_numpy(s), torch.from_numpy(lv),
                             torch.from_numpy(ld), g, lam)
    for n in range(N):
        gae = 0.0
        for t in reversed(range(T)):
            nnt = 1.0 - (ld[n] if t == T - 1 else s[t + 1, n])
            nv = lv[n] if t == T - 1 else v[t + 1, n]
            gae = r[t, n] + g * nv * nnt - v[t, n] + g * lam * nnt * gae
            assert abs(float(adv[t, n]) - gae) < 1e-4
    assert torch.allclose(ret, adv + torch.from_numpy(v))


def make_model(env, **kw):
    args = dict(policy_kwargs=dict(net_arch=dict(pi=[32, 32], vf=[32, 32]), lstm_hidden_size=32), n_steps=48, batch_size=96,
                seq_len=16, n_epochs=2, ent_coef=0.01, seed=1, ops=TorchOps(5))
    args.update(kw)
    return RecurrentPPO(env, **args)


def test_chunk_view_maps_sequence_k_of_env_n_to_column_kN_plus_n():
    env = OracleBatchedEnv(tiny_rooms(), 4, seed=2)
    m = make_model(env)
    T, N, S = 48, 4, 16
    x = torch.arange(T * N).reshape(T, N)
    y = m._chunk_view(x)
    assert y.shape == (S, 3 * N)
    for k in range(3):
        for n in range(N):
            assert torch.equal(y[:, k * N + n], x[k * S:(k + 1) * S, n])


def test_rollout_bookkeeping_and_time_limit_bootstrap():
    env = OracleBatchedEnv(tiny_rooms(), 6, seed=4)
    m = make_model(env)
    const = 10.0
    m.policy.values_step = lambda obs, state, starts: torch.full((obs.shape[0],), const)     # makes the bootstrap visible
    # shadow env with the same seed: replays the actions to get the raw rewards / flags
    shadow = OracleBatchedEnv(tiny_rooms(), 6, seed=4)
    obs0 = shadow.reset().clone()
    assert torch.equal(m._obs[0], obs0)
    assert m.collect_rollouts()
    T = m.n_steps
    assert m.num_timesteps == T * 6
    assert bool(m._starts[0].all())                                   # the very first step of training starts episodes
    n_tl = 0
    for t in range(T):
        _, rew, dones, info = shadow.step(m._actions[t])
        tl = (info.truncated != 0) & (info.terminated == 0)
        n_tl += int(tl.sum())
        assert torch.allclose(m._rewards[t], rew + m.gamma * const * tl.float(), atol=1e-6), t
        assert torch.equal(m._obs[t + 1], shadow._obs)                # env wrote straight into the rollout tensor
        if t + 1 < T:
            assert torch.equal(m._starts[t + 1].bool(), dones)
    assert float(m._chunk_states[0].abs().max()) == 0.0              # the first chunk of training starts from zeros
    assert n_tl > 0, "the tiny rooms must produce time-limit truncations inside one rollout"
    # a second rollout continues where the first stopped (obs[T] -> obs[0], carried LSTM state and episode starts)
    last_obs, last_starts = m._obs[T].clone(), m._last_starts.clone()
    adv1 = m._adv.clone()
    assert m.collect_rollouts()
    assert torch.equal(m._obs[0], last_obs) and torch.equal(m._starts[0], last_starts)
    for t in range(T):
        _, rew, dones, info = shadow.step(m._actions[t])
        assert torch.equal(m._obs[t + 1], shadow._obs)
    assert float(m._chunk_states[0].abs().max()) > 0.0               # chunk 0 now starts from the carried state
    assert m.num_timesteps == 2 * T * 6 and not torch.equal(adv1, m._adv)
    # GAE was computed from those buffers with V(s_T) = const
    adv, ret = gae_reference(m._rewards, m._values, m._starts, torch.full((6,), const), m._last_starts, m.gamma, m.gae_lambda)
    assert torch.allclose(m._adv, adv) and torch.allclose(m._ret, ret)
    assert float(m._chunk_states[1].abs().max()) > 0.0
    assert m._episodes_this_rollout > 0 and math.isfinite(m._ep_return_mean)


def test_learn_save_load_continue(tmp_path):
    env = OracleBatchedEnv(tiny_rooms(), 8, seed=7)
    m = make_model(env, verbose=0)
    before = [p.detach().clone() for p in m.policy.parameters()]
    m.learn(total_timesteps=2 * 48 * 8, reset_num_timesteps=False)
    assert m.num_timesteps == 2 * 48 * 8 and m._iteration == 2 and m.n_updates == 4
    assert any(not torch.equal(a, b) for a, b in zip(before, m.policy.parameters()))
    rec = m.logger[-1]
    for k in ("loss", "policy_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction", "fps"):
        assert math.isfinite(rec[k]), k
    assert rec["minibatches"] == 2 * (3 * 8 // (96 // 16))
    assert -math.log(6) - 1e-3 <= rec["entropy_loss"] < 0
    # gradients still live in the flat buffer (one all-reduce per minibatch in the data-parallel case)
    base = m._flat_grad.data_ptr()
    assert all(base <= p.grad.data_ptr() < base + m._flat_grad.numel() * 4 for p in m.policy.parameters())
    path = m.save(tmp_path / "rppo_hp1_arch_x_lstm_y_s768_view10")
    assert path.name.endswith("_view10.zip")
    odd = m.save(tmp_path / "model_P2.zip_i")                          # Train_Further.py:177's odd suffix is kept
    assert odd.name == "model_P2.zip_i"
    m2 = RecurrentPPO.load(tmp_path / "rppo_hp1_arch_x_lstm_y_s768_view10", env=OracleBatchedEnv(tiny_rooms(), 8, seed=9),
                           ops=TorchOps(6))
    assert m2.num_timesteps == m.num_timesteps and m2.n_updates == 4 and m2.seq_len == 16
    obs = torch.rand(5, 80)
    a1, s1 = m.predict(obs, deterministic=True)
    a2, s2 = m2.predict(obs, deterministic=True)
    assert torch.equal(a1, a2) and all(torch.equal(x, y) for x, y in zip(s1, s2))
    # Adam moments survived the round trip
    st1, st2 = m.optimizer.state_dict()["state"], m2.optimizer.state_dict()["state"]
    assert all(torch.equal(st1[k]["exp_avg"], st2[k]["exp_avg"]) for k in st1)
    m2.learn(total_timesteps=48 * 8, reset_num_timesteps=False)
    assert m2.num_timesteps == 3 * 48 * 8
    # numpy in, numpy out (the form train/evaluate_grid.py uses)
    a, st = m2.predict(np.zeros(80, dtype=np.float32), state=None, episode_start=[True], deterministic=True)
    assert np.ndim(a) == 0 and 0 <= int(a) < 6


def test_evaluate_policy_and_callback(tmp_path):
    env = OracleBatchedEnv(tiny_rooms(), 4, seed=11)
    m = make_model(env)
    eval_env = OracleBatchedEnv(tiny_rooms(), 4, seed=12)
    st = evaluate_policy(m, eval_env, n_eval_episodes=10, deterministic=True, return_episode_stats=True)
    assert len(st["r"]) == 10 and (st["l"] <= 36).all() and (st["l"] >= 1).all()
    mean_r, std_r = evaluate_policy(m, eval_env, n_eval_episodes=6)
    assert math.isfinite(mean_r) and std_r >= 0
    cb = EvalCallback(eval_env, best_model_save_path=tmp_path / "best", log_path=tmp_path / "best", eval_freq=48,
                      n_eval_episodes=10, deterministic=True, verbose=0)
    m.learn(total_timesteps=2 * 48 * 4, callback=cb)
    assert cb.n_calls == 96 and len(cb.evaluations_timesteps) == 2
    assert (tmp_path / "best" / "best_model.zip").exists()
    z = np.load(tmp_path / "best" / "evaluations.npz")
    assert z["results"].shape == (2, 10) and list(z["timesteps"]) == [48 * 4, 96 * 4]


def test_numpy_vec_env_adapter_follows_the_sb3_contract():
    """nav3d.NumpyVecEnv over a batched env: numpy in/out, SB3-style infos on episode ends."""
    from nav3d import NumpyVecEnv
    from nav3d.spaces import cubic_spaces
    inner = OracleBatchedEnv(tiny_rooms(), 6, seed=3)
    inner.action_space, inner.observation_space = cubic_spaces()
    inner.get_attr = lambda name, indices=None: [0] * 6
    inner.env_method = lambda name, *a, indices=None, **k: [None] * 6
    venv = NumpyVecEnv(inner)
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (6, 80) and obs.dtype == np.float32
    rng = np.random.default_rng(0)
    ends = 0
    for t in range(80):
        venv.step_async(rng.integers(0, 6, size=6))
        obs, rew, dones, infos = venv.step_wait()
        assert obs.shape == (6, 80) and rew.dtype == np.float32 and dones.dtype == bool and len(infos) == 6
        for i in range(6):
            if dones[i]:
                ends += 1
                assert infos[i]["terminal_observation"].shape == (80,) and "TimeLimit.truncated" in infos[i]
                assert infos[i]["episode"]["l"] >= 1 and isinstance(infos[i]["episode"]["r"], float)
            else:
                assert infos[i] == {}
    assert ends >= 6                                            # 36- and ~17-step episodes
    with pytest.raises(RuntimeError):
        venv.step_wait()
