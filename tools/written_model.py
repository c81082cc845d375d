#!/usr/bin/env python
"""How many of the 64-byte bricks (4x4 columns x 6 levels) a step reads have never been WRITTEN since the env's reset (and are
therefore still all zero = unknown)?  Cold (steps 5..25), steady (600..650) and late (2000..2100) regimes on real trajectories
of the CPU oracle (random actions, rooms/P1_training, L = 10).  Design tool for a per-env "written bricks" bitmap: a brick
known to be unwritten needs no load.  Test/dev infrastructure; nothing in the product imports it."""
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _nav3d_path  # noqa: E402,F401
from layout_model import marks, L  # noqa: E402


def brick(x, y, z):
    return ((x + 2) >> 2, (y + 2) >> 2, z // 6)


def main():
    from nav3d.rooms import load_room_dir
    from oracle import c_oracle
    rooms = load_room_dir(ROOT / "rooms" / "P1_training", sort=True)
    walls = [r.grid == -2 for r in rooms]
    n, T = 48, 2100
    ov = c_oracle.OracleVec(n, [c_oracle.OracleRoom(r.grid, -2) for r in rooms], L, -2.0, 2024, 0, True)
    ov.reset()
    rng = np.random.default_rng(0)
    prev = ov.state()
    written = [set() for _ in range(n)]

    def first_marks(i, st):
        ri = int(st[i, 13]); x, y, z = (int(v) for v in st[i, :3])
        return {brick(*c) for c in marks(walls[ri], x, y, z, x, y, z, motion=False)} | {brick(x, y, z)}
    for i in range(n):
        written[i] = first_marks(i, prev)
    acc = defaultdict(lambda: defaultdict(float)); cnt = defaultdict(int)
    for t in range(T):
        ov.step(rng.integers(0, 6, size=n))
        st = ov.state()
        for i in range(n):
            ri = int(st[i, 13])
            if st[i, 6] != prev[i, 6] + 1:
                written[i] = first_marks(i, st)
                continue
            sc = int(st[i, 6])
            regime = "cold" if 5 <= sc <= 25 else "steady" if 600 <= sc <= 650 else "late" if 2000 <= sc <= 2100 else None
            x, y, z = (int(v) for v in st[i, :3]); px, py, pz = (int(v) for v in prev[i, :3])
            W, D, H = walls[ri].shape
            first = st[i, 4] == prev[i, 4] + 1
            win = {brick(cx, cy, cz) for cx in range(x - 2, x + 2) for cy in range(y - 2, y + 2) for cz in range(max(z - 2, 0), min(z + 2, H))}
            mk = {brick(*c) for c in marks(walls[ri], x, y, z, px, py, pz) if c != (x, y, z)} if first else set()
            if regime:
                A = acc[regime]; cnt[regime] += 1
                A["win"] += len(win); A["win_unwritten"] += len(win - written[i])
                A["mark"] += len(mk - win); A["mark_unwritten"] += len((mk - win) - written[i])
                A["first"] += first
            written[i] |= mk | {brick(x, y, z)}
        prev = st
    for regime in ("cold", "steady", "late"):
        s = cnt[regime]
        if not s:
            continue
        A = {k: v / s for k, v in acc[regime].items()}
        print(f"{regime}: first-visit rate {A['first']:.2f} | window bricks {A['win']:.2f} (never written {A['win_unwritten']:.2f}) | "
              f"marking bricks outside the window {A['mark']:.2f} (never written {A['mark_unwritten']:.2f}) | "
              f"skippable brick loads per step {A['win_unwritten'] + A['mark_unwritten']:.2f} of {A['win'] + A['mark']:.2f}")


if __name__ == "__main__":
    main()
