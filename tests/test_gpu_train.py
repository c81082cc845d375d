"""-m gpu: the rollout-loop kernels (csrc/nav3d_train.cu, through the C ABI) against their plain-PyTorch fp32
restatements, and the LSTM-PPO trainer end to end on the GPU-resident env."""
import math

import numpy as np
import pytest
import torch

from conftest import ROOMS
from train_refs import gae_reference

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from nav3d.train_ops import DeviceOps
    return DeviceOps(seed=1234, env_id0=0)


@pytest.mark.parametrize("T,N", [(1, 1), (7, 33), (64, 5000), (2048, 8)])
def test_gae_kernel_matches_torch_reference(ops, T, N):
    g = torch.Generator(device="cuda").manual_seed(T * 1000 + N)
    dev = "cuda"
    r = torch.randn((T, N), generator=g, device=dev)
    v = torch.randn((T, N), generator=g, device=dev)
    s = (torch.rand((T, N), generator=g, device=dev) < 0.05).to(torch.uint8)
    lv = torch.randn(N, generator=g, device=dev)
    ld = (torch.rand(N, generator=g, device=dev) < 0.3).to(torch.uint8)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    ops.gae(r, v, s, lv, ld, 0.99, 0.95, adv, ret)
    want_a, want_r = gae_reference(r.cpu(), v.cpu(), s.cpu(), lv.cpu(), ld.cpu(), 0.99, 0.95)
    # fp32 on both sides; the kernel fuses multiply-adds, so allow a few ulp of the running sum (tolerance 1e-5 relative
    # to the largest advantage in the column)
    scale = want_a.abs().max().item() + 1.0
    assert (adv.cpu() - want_a).abs().max().item() <= 1e-5 * scale
    assert (ret.cpu() - want_r).abs().max().item() <= 1e-5 * scale


def test_sample_actions_logprob_entropy_greedy(ops):
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = 3.0 * torch.randn((4099, 6), generator=g, device="cuda")
    ent = torch.empty(4099, device="cuda")
    a, lp = ops.sample_actions(logits, step=3, entropy=ent)
    ref = torch.log_softmax(logits, dim=-1)
    assert a.dtype == torch.int64 and int(a.min()) >= 0 and int(a.max()) <= 5
    assert torch.allclose(lp, ref.gather(-1, a.unsqueeze(-1)).squeeze(-1), atol=2e-6, rtol=1e-5)
    assert torch.allclose(ent, -(ref.exp() * ref).sum(-1), atol=2e-6, rtol=1e-5)
    ga, glp = ops.sample_actions(logits, step=3, greedy=True)
    assert torch.equal(ga, logits.argmax(-1))
    assert torch.allclose(glp, ref.max(-1).values, atol=2e-6, rtol=1e-5)
    # deterministic in (seed, env id, step); a different step gives a different draw
    a2, _ = ops.sample_actions(logits, step=3)
    a3, _ = ops.sample_actions(logits, step=4)
    assert torch.equal(a, a2) and not torch.equal(a, a3)


def test_sample_actions_distribution_and_shard_independence():
    from nav3d.train_ops import DeviceOps
    N = 600_000
    probs = torch.tensor([0.05, 0.4, 0.1, 0.25, 0.15, 0.05])
    logits = probs.log().repeat(N, 1).cuda().contiguous()
    full = DeviceOps(seed=77, env_id0=0)
    a, _ = full.sample_actions(logits, step=11)
    counts = torch.bincount(a, minlength=6).cpu().double()
    expected = probs.double() * N
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    assert chi2 < 25.7, (chi2, counts)                     # 5 dof, p = 1e-4
    # two shards keyed by global env id reproduce the single-shard draw
    half = N // 2
    lo, _ = DeviceOps(seed=77, env_id0=0).sample_actions(logits[:half], step=11)
    hi, _ = DeviceOps(seed=77, env_id0=half).sample_actions(logits[half:], step=11)
    assert torch.equal(torch.cat([lo, hi]), a)
    # successive steps are independent: lag-1 agreement is what independence predicts
    b, _ = full.sample_actions(logits, step=12)
    agree = float((a == b).double().mean())
    assert abs(agree - float((probs ** 2).sum())) < 0.005


def test_train_ops_reject_cpu_tensors(ops):
    with pytest.raises(RuntimeError):
        ops.sample_actions(torch.zeros((4, 6)), step=0)


def test_ppo_learns_on_gpu_env(tmp_path):
    """A short run on an 8x8x6 hollow box (144 free cells, so 144-step episodes): bookkeeping, finite statistics, the
    evaluation callback, a checkpoint round trip — and the policy learns (episode return rises from about -1 for the
    untrained policy to above +20 within 0.8 M steps; measured curve in DESIGN.md)."""
    from nav3d import BatchedCubicEnv
    from nav3d.evaluation import EvalCallback, evaluate_policy
    from nav3d.ppo import RecurrentPPO
    from nav3d.rooms import rooms_from_grids
    g = np.zeros((8, 8, 6), dtype=np.int8)
    g[0], g[-1], g[:, 0], g[:, -1], g[:, :, 0], g[:, :, -1] = -2, -2, -2, -2, -2, -2
    rooms = rooms_from_grids([g])
    env = BatchedCubicEnv(rooms=rooms, num_envs=256, local_map_length=10, seed=3)
    eval_env = BatchedCubicEnv(rooms=rooms, num_envs=8, local_map_length=10, seed=4)
    model = RecurrentPPO(env, policy_kwargs=dict(net_arch=dict(pi=[256, 256, 128], vf=[256, 256, 128]), lstm_hidden_size=256,
                                                 n_lstm_layers=1),
                         learning_rate=3e-4, n_steps=64, batch_size=64 * 64, n_epochs=4, gamma=0.99, gae_lambda=0.95,
                         ent_coef=0.01, vf_coef=0.5, clip_range=0.2, seed=0)
    iters = 48
    cb = EvalCallback(eval_env, best_model_save_path=tmp_path, log_path=tmp_path, eval_freq=64 * 16, n_eval_episodes=8, verbose=0)
    launches0 = env.engine.launch_count
    model.learn(total_timesteps=iters * 64 * 256, callback=cb)
    assert model.num_timesteps == iters * 64 * 256 and len(model.logger) == iters
    assert env.engine.launch_count - launches0 >= 2 * 64           # host-side launches: the eager rollout + the capture; later rollouts are graph replays
    for rec in model.logger:
        assert all(math.isfinite(rec[k]) for k in ("loss", "policy_loss", "value_loss", "entropy_loss", "approx_kl"))
    curve = [r["ep_rew_mean"] for r in model.logger if math.isfinite(r["ep_rew_mean"])]
    assert curve[-1] > curve[0] + 10.0, (curve[0], curve[-1])
    assert len(cb.evaluations_timesteps) == 3 and (tmp_path / "best_model.zip").exists()
    path = model.save(tmp_path / "ckpt")
    again = RecurrentPPO.load(path, env=env)
    obs = torch.rand((16, 80), device=env.device)
    a1, _ = model.predict(obs, deterministic=True)
    a2, _ = again.predict(obs, deterministic=True)
    assert torch.equal(a1, a2)
    st = evaluate_policy(again, eval_env, n_eval_episodes=8, return_episode_stats=True)
    assert len(st["r"]) == 8 and (st["total_free"] == 144).all()


@pytest.mark.parametrize("S,B,F,H,p_start", [(1, 3, 5, 8, 0.5), (9, 6, 7, 10, 0.3), (17, 33, 80, 64, 0.15), (128, 512, 80, 256, 0.01), (64, 700, 80, 256, 0.0)])
def test_fused_lstm_matches_torch_lstm(S, B, F, H, p_start):
    """nav3d_lstm_forward / nav3d_lstm_backward against torch.nn.LSTM stepped one timestep at a time with the state
    zeroed at episode starts (plain fp32, TF32 off on both sides): outputs, final state and all four parameter gradients.
    Tolerance: 2e-5 absolute on outputs in [-1, 1], 1e-4 relative to each gradient's largest entry (fp32 sums over up to
    65 536 rows in a different order)."""
    from nav3d.train_ops import fused_lstm
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda").manual_seed(S * 7 + B)
        lstm = torch.nn.LSTM(F, H).cuda()
        x = torch.randn((S, B, F), generator=g, device="cuda")
        h0 = torch.randn((B, H), generator=g, device="cuda") * 0.5
        c0 = torch.randn((B, H), generator=g, device="cuda") * 0.5
        starts = (torch.rand((S, B), generator=g, device="cuda") < p_start).to(torch.uint8)
        wgt = torch.randn((S, B, H), generator=g, device="cuda")
        # reference
        st = (h0.unsqueeze(0), c0.unsqueeze(0))
        outs = []
        for t in range(S):
            keep = (1.0 - starts[t].float()).view(1, B, 1)
            y, st = lstm(x[t:t + 1], (st[0] * keep, st[1] * keep))
            outs.append(y)
        ref = torch.cat(outs)
        lstm.zero_grad()
        (ref * wgt).sum().backward()
        ref_grads = [p.grad.clone() for p in lstm.parameters()]
        # fused
        lstm.zero_grad()
        y, h_last, c_last = fused_lstm(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, h0, c0, starts, tf32=False)
        assert (y - ref.detach()).abs().max().item() < 2e-5
        assert (h_last - st[0][0].detach()).abs().max().item() < 2e-5 and (c_last - st[1][0].detach()).abs().max().item() < 5e-5
        (y * wgt).sum().backward()
        for p, r in zip(lstm.parameters(), ref_grads):
            assert (p.grad - r).abs().max().item() <= 1e-4 * (r.abs().max().item() + 1e-6), (p.shape, (p.grad - r).abs().max().item(), r.abs().max().item())
        # TF32 tensor-op math for the GEMMs: same function to TF32 accuracy
        y32, _, _ = fused_lstm(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, h0, c0, starts, tf32=True)
        assert (y32 - ref.detach()).abs().max().item() < 2e-2
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_policy_fused_and_cudnn_paths_agree():
    """RecurrentActorCritic.forward_sequence through the fused LSTM (default on CUDA) and through cuDNN with cuts, and with
    the critic branch on its own stream or not: same logits/values and the same gradients (TF32 off)."""
    from nav3d.policy import RecurrentActorCritic
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        pol = RecurrentActorCritic(net_arch=dict(pi=[256, 256, 128], vf=[256, 256, 128]), lstm_hidden_size=256).cuda()
        S, B = 32, 96
        obs = torch.rand((S, B, 80), device="cuda")
        starts = (torch.rand((S, B), device="cuda") < 0.05).to(torch.uint8)
        state = tuple(0.3 * torch.randn((1, B, 256), device="cuda") for _ in range(4))
        cuts = [0] + [t for t in range(1, S) if bool(starts[t].any())]
        results = []
        for fused, two in ((False, False), (True, False), (True, True), (False, True)):
            pol.fused_lstm, pol.two_streams = fused, two
            pol.zero_grad()
            logits, values, st = pol.forward_sequence(obs, state, starts, cuts)
            (logits.square().mean() + values.square().mean()).backward()
            torch.cuda.synchronize()
            results.append((logits.detach(), values.detach(), [s.detach() for s in st], [p.grad.clone() for p in pol.parameters()]))
        base = results[0]
        for r in results[1:]:
            assert torch.allclose(r[0], base[0], atol=1e-5) and torch.allclose(r[1], base[1], atol=1e-5)
            assert all(torch.allclose(a, b, atol=5e-5) for a, b in zip(r[2], base[2]))
            for a, b in zip(r[3], base[3]):
                assert (a - b).abs().max().item() <= 1e-4 * (b.abs().max().item() + 1e-6)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_cuda_graph_rollout_equals_eager_rollout():
    """From the second rollout on, collect_rollouts replays ONE CUDA graph of the whole rollout; it must fill the buffers
    exactly like the eager loop (same Philox draws through the device-side step counter, same env steps, same GAE)."""
    from nav3d import BatchedCubicEnv
    from nav3d.ppo import RecurrentPPO
    from nav3d.rooms import load_room_file
    # 149- and 302-step episodes: resets and time limits fall inside every rollout after the first
    rooms = [load_room_file(ROOMS / "P3_training" / "maze_3d_tunnels.txt"), load_room_file(ROOMS / "P2_training" / "tightcorridor.txt")]
    models = []
    for graph in (False, True):
        env = BatchedCubicEnv(rooms=rooms, num_envs=192, local_map_length=10, seed=5)
        m = RecurrentPPO(env, policy_kwargs=dict(net_arch=dict(pi=[64, 64], vf=[64, 64]), lstm_hidden_size=64), n_steps=96,
                         batch_size=96 * 64, n_epochs=1, seed=3, cuda_graph=graph, graph_chunk=40)   # 3 graphs: 40 + 40 + 16 steps
        models.append(m)
    calls = [0, 0]

    class Count:
        def __init__(self, i): self.i = i
        def on_step(self, model):
            calls[self.i] += 1
            return True
    for r in range(4):
        for i, m in enumerate(models):
            assert m.collect_rollouts(Count(i))
        a, b = models
        assert (b._graph is not None) == (r >= 1) and a._graph is None and (r < 1 or len(b._graph) == 3)
        assert torch.equal(a._actions, b._actions) and torch.equal(a._obs, b._obs) and torch.equal(a._starts, b._starts)
        assert torch.allclose(a._rewards, b._rewards, atol=1e-5) and torch.allclose(a._values, b._values, atol=1e-5)
        assert torch.allclose(a._adv, b._adv, atol=1e-4) and torch.allclose(a._logp, b._logp, atol=1e-5)
        assert torch.allclose(a._chunk_states, b._chunk_states, atol=1e-5)
        assert a.num_timesteps == b.num_timesteps == (r + 1) * 96 * 192 and calls[0] == calls[1] == (r + 1) * 96
        assert a._episodes_this_rollout == b._episodes_this_rollout
        assert (math.isnan(a._ep_return_mean) and math.isnan(b._ep_return_mean)) or abs(a._ep_return_mean - b._ep_return_mean) < 1e-3
    assert int(a._starts.sum()) > 0 and a._episodes_this_rollout > 0
    # training between graph replays keeps working (the graph reads the parameters in place)
    b.train()
    assert b.collect_rollouts()


def test_cuda_graph_update_equals_eager_update():
    """Small minibatches (launch-bound) are replayed as ONE CUDA graph per minibatch (forward, backward, clip, Adam); after
    the same rollouts and the same shuffles the parameters must agree with the eager update."""
    from nav3d import BatchedCubicEnv
    from nav3d.ppo import RecurrentPPO
    from nav3d.rooms import load_room_file
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    rooms = [load_room_file(ROOMS / "P3_training" / "maze_3d_tunnels.txt")]
    try:
        models = []
        for graph in (False, True):
            env = BatchedCubicEnv(rooms=rooms, num_envs=8, local_map_length=10, seed=5)
            m = RecurrentPPO(env, policy_kwargs=dict(net_arch=dict(pi=[64, 64], vf=[64, 64]), lstm_hidden_size=64), n_steps=64,
                             batch_size=32, n_epochs=2, ent_coef=0.01, seed=3, cuda_graph=graph, allow_tf32=False)
            models.append(m)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        a, b = models
        # two iterations: the first holds the eager warm-up minibatches, the capture and replays, the second replays only.
        # (Adam turns rounding-level gradient differences into lr-sized steps, so longer runs drift apart legitimately.)
        for it in range(2):
            for m in models:
                m.collect_rollouts()
            assert torch.equal(a._actions, b._actions), it            # same policy so far -> same rollout
            sa, sb = a.train(), b.train()
            assert sa["minibatches"] == sb["minibatches"] == 2 * (64 * 8 // 32)
            pa = torch.cat([p.detach().flatten() for p in a.policy.parameters()])
            pb = torch.cat([p.detach().flatten() for p in b.policy.parameters()])
            rel = float((pa - pb).norm() / pa.norm())
            assert rel < 5e-4, (it, rel)
            for k in ("policy_loss", "value_loss", "entropy_loss", "approx_kl"):
                assert abs(sa[k] - sb[k]) <= 1e-3 * (1.0 + abs(sa[k])), (k, sa[k], sb[k])
        assert b._upd_graph is not None and a._upd_graph is None
        # the optimizer state of a graph-trained model survives a checkpoint round trip
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            path = b.save(d + "/m")
            c = RecurrentPPO.load(path, env=BatchedCubicEnv(rooms=rooms, num_envs=8, local_map_length=10, seed=5))
            c.collect_rollouts(); c.train()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_cuda_graph_update_masks_episode_starts_inside_a_chunk():
    """episodes that end in the middle of a sequence chunk.  A 5x6x5 room has 36 free cells, so every
    episode is truncated after 36 steps and each 64-step chunk holds one or two resets; the replayed (graph) update
    must zero the LSTM state there exactly like the eager update: same losses, same parameters."""
    import numpy as np
    from edge_rooms import rooms_from_grids
    from nav3d import BatchedCubicEnv
    from nav3d.ppo import RecurrentPPO
    g = np.zeros((5, 6, 5), dtype=np.int8)
    g[0], g[-1], g[:, 0], g[:, -1], g[:, :, 0], g[:, :, -1] = -2, -2, -2, -2, -2, -2
    rooms = rooms_from_grids([g])
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        models = []
        for graph in (False, True):
            env = BatchedCubicEnv(rooms=rooms, num_envs=8, local_map_length=10, seed=5)
            models.append(RecurrentPPO(env, policy_kwargs=dict(net_arch=dict(pi=[64, 64], vf=[64, 64]), lstm_hidden_size=64),
                                       n_steps=64, batch_size=64, n_epochs=2, ent_coef=0.01, seed=3, cuda_graph=graph,
                                       allow_tf32=False))
        a, b = models
        for it in range(2):
            for m in models:
                m.collect_rollouts()
            assert torch.equal(a._actions, b._actions), it
            starts = a._starts.view(64, 8)
            assert int(starts[1:].sum()) >= 8                      # resets strictly inside the chunks
            sa, sb = a.train(), b.train()
            pa = torch.cat([p.detach().flatten() for p in a.policy.parameters()])
            pb = torch.cat([p.detach().flatten() for p in b.policy.parameters()])
            rel = float((pa - pb).norm() / pa.norm())
            assert rel < 5e-4, (it, rel)
            for k in ("policy_loss", "value_loss", "entropy_loss", "approx_kl"):
                assert abs(sa[k] - sb[k]) <= 1e-3 * (1.0 + abs(sa[k])), (k, sa[k], sb[k])
        assert b._upd_graph is not None and a._upd_graph is None
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
