import sys, torch
sys.path.insert(0, "."); import _nav3d_path
from torch.profiler import profile, ProfilerActivity
from nav3d import BatchedCubicEnv
from nav3d.ppo import RecurrentPPO
env = BatchedCubicEnv("rooms/P1_training", num_envs=4096, local_map_length=10, seed=42, sort_rooms=True)
m = RecurrentPPO(env, policy_kwargs=dict(net_arch=dict(pi=[256, 256, 128], vf=[256, 256, 128]), lstm_hidden_size=256, n_lstm_layers=1),
                 n_steps=128, batch_size=2048 * 128, n_epochs=2, ent_coef=0.01, seed=0)
m.policy.two_streams = False
m.collect_rollouts(); m.train(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.train(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.collect_rollouts(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
