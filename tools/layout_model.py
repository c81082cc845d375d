#!/usr/bin/env python
"""Distinct 64-byte DRAM atoms one env-step touches in the per-env knowledge block, for several candidate layouts and three
regimes (cold = steps 5..25 after a reset, the regime `bench.py --steps 20 --warmup 5` times; warm = steps 50..1000;
late = steps 2000..3000), on real trajectories of the CPU oracle (random actions, rooms/P1_training, L = 10).

This is the design tool behind DESIGN.md §6: the per-step kernel is bound by scattered 64-byte DRAM atoms, so a layout is
judged by the atoms it touches per step.  Test/dev infrastructure (drives oracle/); nothing in the product imports it.
usage: python tools/layout_model.py [--envs N] [--steps T]"""
import argparse
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _nav3d_path  # noqa: E402,F401

L = 10


def ray_extent(wall_line, p):
    n = len(wall_line)
    up = dn = 0
    for s in range(1, L + 1):
        if p + s >= n:
            break
        up = s
        if wall_line[p + s]:
            break
    for s in range(1, L + 1):
        if p - s < 0:
            break
        dn = s
        if wall_line[p - s]:
            break
    return dn, up


def marks(wl, x, y, z, px, py, pz, motion=True):
    """cells whose seen state a first visit at (x,y,z) coming from (px,py,pz) may change: list of (cx, cy, cz)"""
    xd, xu = ray_extent(wl[:, y, z], x)
    yd, yu = ray_extent(wl[x, :, z], y)
    zd, zu = ray_extent(wl[x, y, :], z)
    xs = list(range(x - xd, x + xu + 1))
    ys = list(range(y - yd, y + yu + 1))
    zs = list(range(z - zd, z + zu + 1))
    if motion and (x != px):
        far = x + L if x > px else x - L
        xs = [x] + ([far] if ((xu == L and x > px) or (xd == L and x < px)) else [])
    elif motion and (y != py):
        far = y + L if y > py else y - L
        ys = [y] + ([far] if ((yu == L and y > py) or (yd == L and y < py)) else [])
    elif motion and (z != pz):
        far = z + L if z > pz else z - L
        zs = [z] + ([far] if ((zu == L and z > pz) or (zd == L and z < pz)) else [])
    out = {(cx, y, z) for cx in xs} | {(x, cy, z) for cy in ys} | {(x, y, cz) for cz in zs}
    return out


class Dims:
    def __init__(self, W, D, H):
        self.W, self.D, self.H = W, D, H
        self.ntx, self.nty = (W + 3) // 4, (D + 3) // 4


# ---- layouts: each maps cells -> atom ids (ints, unique per structure via a tag) ---------------------------------------
def s_tile_atom(d, x, y):                       # shipped S: 4x4 columns x u16 = 32 B, x-adjacent tiles share an atom
    return ("S", ((y >> 2) * d.ntx + (x >> 2)) >> 1)


def c_brick_atom(d, x, y, z):                   # shipped C: 4x4x2 u8 = 32 B, z-adjacent bricks share an atom
    nbz = (d.H + 1) // 2
    return ("C", (((y >> 2) * d.ntx + (x >> 2)) * nbz + (z >> 1)) >> 1)


def c444_atom(d, x, y, z):                      # 4x4x4 bytes = 64 B
    return ("B", (x >> 2), (y >> 2), (z >> 2))


def sv_tile_atom(d, x, y):                      # S+V tile: 4x4 columns x (u16 seen + u16 visited) = 64 B
    return ("SV", (x >> 2), (y >> 2))


def col64_atom(tx, ty):
    def f(d, x, y):                             # u64 code column (all z), tx x ty columns per atom
        return ("K", x // tx, y // ty)
    return f


def nib_atom(bx, by, bz):
    def f(d, x, y, z):                          # 4-bit cells, bx*by*bz = 128 cells per atom
        return ("N", x // bx, y // by, z // bz)
    return f


def window_cells(d, x, y, z):
    return [(cx, cy, cz) for cx in range(x - 2, x + 2) for cy in range(y - 2, y + 2) for cz in range(z - 2, z + 2)
            if 0 <= cx < d.W and 0 <= cy < d.D and 0 <= cz < d.H]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3000)
    args = ap.parse_args()
    from nav3d.rooms import load_room_dir
    from oracle import c_oracle
    rooms = load_room_dir(ROOT / "rooms" / "P1_training", sort=True)
    walls = [r.grid == -2 for r in rooms]
    dims = [Dims(*r.grid.shape) for r in rooms]
    n = args.envs
    ov = c_oracle.OracleVec(n, [c_oracle.OracleRoom(r.grid, -2) for r in rooms], L, -2.0, 2024, 0, True)
    ov.reset()
    rng = np.random.default_rng(0)
    prev = ov.state()
    visited = [np.zeros(walls[int(prev[i, 13])].shape, bool) for i in range(n)]
    for i in range(n):
        visited[i][tuple(prev[i, :3])] = True
    pending = [[] for _ in range(n)]            # deferred marking: first-visit entries not yet written
    bbox = [[*prev[i, :3], *prev[i, :3]] for i in range(n)]
    acc = defaultdict(lambda: defaultdict(float))
    cnt = defaultdict(int)
    KDEF = (4, 8)
    pend = {k: [[] for _ in range(n)] for k in KDEF}
    for t in range(args.steps):
        a = rng.integers(0, 6, size=n)
        ov.step(a)
        st = ov.state()
        for i in range(n):
            ri = int(st[i, 13])
            if st[i, 6] != prev[i, 6] + 1:       # auto-reset in this step: restart the bookkeeping
                visited[i] = np.zeros(walls[ri].shape, bool)
                visited[i][tuple(st[i, :3])] = True
                for k in KDEF:
                    pend[k][i] = []
                continue
            sc = int(st[i, 6])
            regime = "cold" if 5 <= sc <= 25 else "warm" if 50 <= sc <= 1000 else "late" if 2000 <= sc <= 3000 else None
            wl, d = walls[ri], dims[ri]
            x, y, z = (int(v) for v in st[i, :3])
            px, py, pz = (int(v) for v in prev[i, :3])
            first = st[i, 4] == prev[i, 4] + 1
            V = visited[i]
            V[x, y, z] = True
            wc = window_cells(d, x, y, z)
            mk = marks(wl, x, y, z, px, py, pz) if first else set()
            vis_w = [c for c in wc if V[c]]
            if regime is None:
                for k in KDEF:                   # keep the deferred lists moving outside the sampled regimes too
                    if first:
                        pend[k][i].append((x, y, z, px, py, pz))
                        if len(pend[k][i]) >= k:
                            pend[k][i] = []
                continue
            cnt[regime] += 1
            A = acc[regime]
            A["first"] += first
            A["visited_in_window"] += len(vis_w)
            # -- shipped: S tiles + C bricks
            sa = {s_tile_atom(d, cx, cy) for cx, cy, _ in wc}
            ca = {c_brick_atom(d, *c) for c in wc}
            ma = {s_tile_atom(d, cx, cy) for cx, cy, _ in mk}
            A["ship:S"] += len(sa); A["ship:C"] += len(ca); A["ship:mark"] += len(ma - sa)
            # -- fold (seen bit in the counter byte), 4x4x4 bricks
            ba = {c444_atom(d, *c) for c in wc}
            mb = {c444_atom(d, *c) for c in mk}
            A["fold:win"] += len(ba); A["fold:mark"] += len(mb - ba)
            # -- S+V tiles (64 B per 4x4 columns) + counter bricks only where a visited window cell lies
            sva = {sv_tile_atom(d, cx, cy) for cx, cy, _ in wc}
            cva = {c444_atom(d, *c) for c in vis_w}
            mva = {sv_tile_atom(d, cx, cy) for cx, cy, _ in mk}
            A["sv:SV"] += len(sva); A["sv:C"] += len(cva); A["sv:mark"] += len(mva - sva)
            # -- shipped S tiles + V-gated shipped C bricks (V kept in a separate, S-shaped bit volume)
            va = {("V",) + s_tile_atom(d, cx, cy)[1:] for cx, cy, _ in wc}
            cga = {c_brick_atom(d, *c) for c in vis_w}
            A["sgate:S+V"] += len(sa) + len(va); A["sgate:C"] += len(cga); A["sgate:mark"] += len(ma - sa)
            # -- u64 code columns
            for name, (tx, ty) in (("k42", (4, 2)), ("k24", (2, 4))):
                f = col64_atom(tx, ty)
                ka = {f(d, cx, cy) for cx, cy, _ in wc}
                km = {f(d, cx, cy) for cx, cy, _ in mk}
                A[name + ":win"] += len(ka); A[name + ":mark"] += len(km - ka)
            # -- nibble bricks
            for name, b in (("n448", (4, 4, 8)), ("n844", (8, 4, 4)), ("n484", (4, 8, 4)), ("k6_446", (4, 4, 6)), ("k6_826", (8, 2, 6)), ("k6_843", (8, 4, 3))):
                f = nib_atom(*b)
                na = {f(d, *c) for c in wc}
                nm = {f(d, *c) for c in mk}
                A[name + ":win"] += len(na); A[name + ":mark"] += len(nm - na)
            # -- deferred marking on the shipped S tiles: flush K first visits at once
            for k in KDEF:
                if first:
                    pend[k][i].append((x, y, z, px, py, pz))
                    if len(pend[k][i]) >= k:
                        allm = set()
                        for (qx, qy, qz, rx, ry, rz) in pend[k][i]:
                            allm |= {s_tile_atom(d, cx, cy) for cx, cy, _ in marks(wl, qx, qy, qz, rx, ry, rz)}
                        A[f"defer{k}:mark"] += len(allm - sa)
                        pend[k][i] = []
        prev = st
    for regime in ("cold", "warm", "late"):
        s = cnt[regime]
        if not s:
            continue
        A = {k: v / s for k, v in acc[regime].items()}
        print(f"== {regime}: {s} env-steps, first-visit rate {A['first']:.3f}, visited cells in window {A['visited_in_window']:.1f}")
        print(f"  shipped (S tiles + C bricks)      : S {A['ship:S']:.2f} + C {A['ship:C']:.2f} + mark {A['ship:mark']:.2f} = "
              f"{A['ship:S'] + A['ship:C'] + A['ship:mark']:.2f}")
        for k in KDEF:
            print(f"  shipped + deferred marking K={k}     : S {A['ship:S']:.2f} + C {A['ship:C']:.2f} + mark {A[f'defer{k}:mark']:.2f} = "
                  f"{A['ship:S'] + A['ship:C'] + A[f'defer{k}:mark']:.2f}")
        print(f"  fold seen into 4x4x4 byte bricks  : win {A['fold:win']:.2f} + mark {A['fold:mark']:.2f} = {A['fold:win'] + A['fold:mark']:.2f}")
        print(f"  S+V 64-B tiles, V-gated 4x4x4 C   : SV {A['sv:SV']:.2f} + C {A['sv:C']:.2f} + mark {A['sv:mark']:.2f} = "
              f"{A['sv:SV'] + A['sv:C'] + A['sv:mark']:.2f}")
        print(f"  S tiles + V tiles, V-gated C      : S+V {A['sgate:S+V']:.2f} + C {A['sgate:C']:.2f} + mark {A['sgate:mark']:.2f} = "
              f"{A['sgate:S+V'] + A['sgate:C'] + A['sgate:mark']:.2f}")
        for name in ("k42", "k24", "n448", "n844", "n484", "k6_446", "k6_826", "k6_843"):
            print(f"  {name:34s}: win {A[name + ':win']:.2f} + mark {A[name + ':mark']:.2f} = {A[name + ':win'] + A[name + ':mark']:.2f}")


if __name__ == "__main__":
    main()
