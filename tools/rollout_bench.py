#!/usr/bin/env python
"""Fused random-action rollout (nav3d_rollout_random) throughput for the output sets a caller may ask for.
usage: python tools/rollout_bench.py [--envs N --T T --lanes G]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _nav3d_path  # noqa: E402,F401


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=32)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--only", default="", help="run only the variant whose name contains this")
    ap.add_argument("--preroll", type=int, default=0, help="untimed fused steps before the measurement (episode age)")
    args = ap.parse_args()
    import torch
    from nav3d import Engine
    from nav3d.rooms import load_room_dir
    rooms = load_room_dir(ROOT / "rooms" / "P1_training", sort=True)
    n, T = args.envs, args.T
    eng = Engine(n, rooms, local_map_length=10, seed=2024, lanes_per_env=args.lanes)
    dev = eng.device
    obs_last = eng.reset()
    rew = torch.empty((T, n), dtype=torch.float32, device=dev)
    done = torch.empty((T, n), dtype=torch.uint8, device=dev)
    out = {"envs": n, "T": T, "lanes": eng.lanes_per_env}
    variants = {"last_obs_only": dict(obs_last=obs_last), "last_obs+reward+done": dict(obs_last=obs_last, reward=rew, done=done)}
    if n * T * 320 < 40e9:
        variants["all_obs+reward+done"] = dict(obs=torch.empty((T, n, 80), dtype=torch.float32, device=dev), reward=rew, done=done)
    t0 = 0
    if args.preroll:
        for _ in range(args.preroll // T):
            eng.rollout_random(T, t0, obs_last=obs_last); t0 += T
        out["preroll"] = t0
    for name, kw in variants.items():
        if args.only and args.only not in name:
            continue
        eng.rollout_random(T, t0, **kw); t0 += T
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            eng.rollout_random(T, t0, **kw); t0 += T
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"env_steps_per_s": n * T * args.reps / (ms / 1e3), "ms_per_env_step_batch": ms / (T * args.reps)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
