#!/usr/bin/env python
"""Per-launch durations of nav3d_step (CUDA events around every launch) after a reset and after a fused pre-roll:
shows how the step time evolves with the age of the episodes.  usage: python tools/per_launch.py [--envs N]"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _nav3d_path  # noqa: E402,F401


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=160)
    ap.add_argument("--preroll", type=int, default=600)
    args = ap.parse_args()
    import torch
    from nav3d import Engine
    from nav3d.rooms import load_room_dir
    rooms = load_room_dir(ROOT / "rooms" / "P1_training", sort=True)
    n = args.envs
    eng = Engine(n, rooms, local_map_length=10, seed=2024)
    dev = eng.device
    g = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, 6, (16, n), generator=g, device=dev, dtype=torch.int64)
    obs = torch.empty((4, n, 80), device=dev); rew = torch.empty(n, device=dev)
    te = torch.empty(n, dtype=torch.uint8, device=dev); tr = torch.empty(n, dtype=torch.uint8, device=dev)

    def run(k):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        torch.cuda.synchronize()
        ev[0].record()
        for t in range(k):
            eng.step(acts[t % 16], obs[t % 4], rew, te, tr)
            ev[t + 1].record()
        torch.cuda.synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(k)]

    eng.reset(obs[0])
    a = run(args.steps)
    for i in range(args.preroll // 40):
        eng.rollout_random(40, 10_000 + 40 * i, obs_last=obs[0])
    b = run(args.steps)
    fmt = lambda v: " ".join(f"{x:.0f}" for x in v)
    print("after reset   (us per launch, every 8th):", fmt(a[::8]))
    print("after preroll (us per launch, every 8th):", fmt(b[::8]))


if __name__ == "__main__":
    main()
