#!/usr/bin/env python
"""Count the distinct 64-byte DRAM atoms one env-step reads from the per-env knowledge block, for the shipped layout and for
the alternatives DESIGN.md §6 discusses — on real trajectories (CPU oracle, random actions, rooms/P1_training).

The step kernel is bound by scattered 64-byte reads, so atoms per step is the quantity a layout has to
reduce.  The model replays what the kernel touches: the S tile and the C bricks of every in-bounds column of the 4x4x4
window, and on a first visit the S tiles of the (motion-aware) ray marking.  Usage: python tools/atom_model.py [--envs N --steps T]"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _nav3d_path  # noqa: E402,F401


def ray_extent(wall_line, p, L):
    """cells p+1 .. along +, and p-1 .. along -, up to and including the first wall, at most L, clipped to the room."""
    n = len(wall_line)
    up = 0
    for s in range(1, L + 1):
        if p + s >= n:
            break
        up = s
        if wall_line[p + s]:
            break
    dn = 0
    for s in range(1, L + 1):
        if p - s < 0:
            break
        dn = s
        if wall_line[p - s]:
            break
    return dn, up


class Layout:
    """Byte offsets inside one env's knowledge block, as in csrc/nav3d_core.cuh (s_index / c_index)."""

    def __init__(self, W, D, H):
        self.ntx, self.nty, self.nbz = (W + 3) // 4, (D + 3) // 4, (H + 1) // 2
        self.c_off = -(-self.ntx * self.nty * 32 // 128) * 128

    def s_atom(self, x, y):
        return (((y >> 2) * self.ntx + (x >> 2)) * 32) // 64

    def c_atom(self, x, y, zb):
        return (self.c_off + (((y >> 2) * self.ntx + (x >> 2)) * self.nbz + zb) * 32) // 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=96)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--skip", type=int, default=50)
    args = ap.parse_args()
    from nav3d.rooms import load_room_dir
    from oracle import c_oracle
    rooms = load_room_dir(ROOT / "rooms" / "P1_training", sort=True)
    walls = [r.grid == -2 for r in rooms]
    lay = [Layout(*r.grid.shape) for r in rooms]
    L = 10
    n = args.envs
    ov = c_oracle.OracleVec(n, [c_oracle.OracleRoom(r.grid, -2) for r in rooms], L, -2.0, 2024, 0, True)
    ov.reset()
    rng = np.random.default_rng(0)
    prev = ov.state()
    tot = dict(window_S=0.0, window_C=0.0, marking=0.0, marking_full=0.0, face_S=0.0, face_C=0.0, steps=0, first=0, moved=0)
    for t in range(args.steps):
        a = rng.integers(0, 6, size=n)
        ov.step(a)
        st = ov.state()
        if t >= args.skip:
            for i in range(n):
                if st[i, 6] != prev[i, 6] + 1:           # reset happened in this step: skip
                    continue
                ri = int(st[i, 13])
                wl, ly = walls[ri], lay[ri]
                W, D, H = wl.shape
                x, y, z = (int(v) for v in st[i, :3])
                px, py, pz = (int(v) for v in prev[i, :3])
                moved = (x, y, z) != (px, py, pz)
                first = st[i, 4] == prev[i, 4] + 1
                zb0 = (z - 2) >> 1
                zbs = [zb for zb in (zb0, zb0 + 1) + ((zb0 + 2,) if (z - 2) & 1 else ()) if 0 <= zb < ly.nbz]
                sa, ca = set(), set()
                for cx in range(x - 2, x + 2):
                    for cy in range(y - 2, y + 2):
                        if 0 <= cx < W and 0 <= cy < D:
                            sa.add(ly.s_atom(cx, cy))
                            for zb in zbs:
                                ca.add(ly.c_atom(cx, cy, zb))
                tot["window_S"] += len(sa)
                tot["window_C"] += len(ca)
                tot["steps"] += 1
                tot["moved"] += moved
                if moved:                                  # entering face of the window (for the cache variants)
                    fs, fc = set(), set()
                    dx, dy, dz = x - px, y - py, z - pz
                    for cx in range(x - 2, x + 2):
                        for cy in range(y - 2, y + 2):
                            if not (0 <= cx < W and 0 <= cy < D):
                                continue
                            new_col = (dx and cx == (x + 1 if dx > 0 else x - 2)) or (dy and cy == (y + 1 if dy > 0 else y - 2))
                            if new_col:
                                fs.add(ly.s_atom(cx, cy))
                                for zb in zbs:
                                    fc.add(ly.c_atom(cx, cy, zb))
                            elif dz:
                                zn = z + 1 if dz > 0 else z - 2
                                if 0 <= zn < H:
                                    fc.add(ly.c_atom(cx, cy, zn >> 1))
                    tot["face_S"] += len(fs)
                    tot["face_C"] += len(fc)
                if first:
                    tot["first"] += 1
                    xd, xu = ray_extent(wl[:, y, z], x, L)
                    yd, yu = ray_extent(wl[x, :, z], y, L)
                    full = {ly.s_atom(cx, y) for cx in range(x - xd, x + xu + 1)} | {ly.s_atom(x, cy) for cy in range(y - yd, y + yu + 1)}
                    tot["marking_full"] += len(full)
                    m = {ly.s_atom(x, y)}
                    if x != px:
                        m |= {ly.s_atom(x, cy) for cy in range(y - yd, y + yu + 1)}
                        far = x + L if x > px else x - L
                        if (xu == L and x > px) or (xd == L and x < px):
                            m.add(ly.s_atom(far, y))
                    elif y != py:
                        m |= {ly.s_atom(cx, y) for cx in range(x - xd, x + xu + 1)}
                        far = y + L if y > py else y - L
                        if (yu == L and y > py) or (yd == L and y < py):
                            m.add(ly.s_atom(x, far))
                    else:
                        m = full
                    tot["marking"] += len(m - sa)          # atoms not already fetched for the window
        prev = st
    s = tot["steps"]
    ws, wc, mk, mf = tot["window_S"] / s, tot["window_C"] / s, tot["marking"] / s, tot["marking_full"] / s
    fS, fC = tot["face_S"] / s, tot["face_C"] / s
    print(f"steps {s}  first-visit rate {tot['first'] / s:.3f}  move rate {tot['moved'] / s:.3f}")
    print(f"shipped layout       : window S {ws:.2f} + window C {wc:.2f} + ray marking {mk:.2f} (without motion-awareness {mf:.2f}) "
          f"= {ws + wc + mk:.2f} scattered atoms per step (+ record, actions: streamed)")
    print(f"S-column cache       : entering S columns {fS:.2f} + window C {wc:.2f} + marking {mk:.2f} = {fS + wc + mk:.2f}  (+64 B streamed record)")
    print(f"window cache (S + C) : entering faces S {fS:.2f} + C {fC:.2f} + counter write-through 1 + marking {mk:.2f} = "
          f"{fS + fC + 1 + mk:.2f}  (+192 B streamed record)")


if __name__ == "__main__":
    main()
