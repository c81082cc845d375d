"""Tools around the reference's ``rooms/*.txt`` voxel format: a writer that is the exact inverse of
the parser, a validator for the quirks that silently change an environment, and a small procedural generator.

The format and its parser are the reference's (``README.md:8-17``; ``envs/CubicEnv.py:402-438``, restated in
``nav3d/rooms.py``).  The reference ships no tooling of its own for these files; the checks below come from what its loader
silently accepts: ``Layer z=-k`` wraps to ``h-k`` (``rooms/P3_training/kitchen2.txt:62``); cells
other than 0/2 are kept as free cells with odd values; a boundary shell that is not all wall lets the agent reach the room
edge (real out-of-bounds paths in ``maze_7x7_*``, ``maze_8x8_*``, ``maze_dead1_11``, ``kitchen2``); free cells that cannot
be reached from each other make the 84 % finish line (``CubicEnv.py:12``, ``:212``) unreachable.

    python -m nav3d.room_tools validate rooms/P3_training
    python -m nav3d.room_tools generate --kind maze --size 21,21,9 --seed 3 -o my_maze.txt
    python -m nav3d.room_tools compose --objects rooms/objects --size 32,32,12 --count 6 -o my_flat.txt"""
from __future__ import annotations

import argparse
import sys
from collections import deque
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

from .rooms import CUBIC_WALL, Room, parse_room_text

MAX_W, MAX_D, MAX_H = 64, 64, 16          # include/nav3d.h limits


# ---- writer ----------------------------------------------------------------------------------------------------------
def room_to_text(grid: np.ndarray, start: Optional[Tuple[int, int, int]] = None,
                 goal: Optional[Tuple[int, int, int]] = None, *, cubic: bool = True) -> str:
    """Serialise ``grid[x, y, z]`` so that ``parse_room_text`` gives it back.  ``cubic``: the grid uses CubicEnv's codes
    (-2 = wall) and walls are written as the file token ``2``; otherwise values are written as they are."""
    g = np.asarray(grid)
    w, d, h = g.shape
    out = [f"Size={w},{d},{h}"]
    for z in range(h):
        out.append(f"Layer z={z}")
        for y in range(d):
            vals = g[:, y, z]
            if cubic:
                vals = [2 if v == CUBIC_WALL else int(v) for v in vals]
            out.append(" ".join(str(int(v)) for v in vals))
    if start is not None:
        out.append("Start position=%d,%d,%d" % tuple(start))
    if goal is not None:
        out.append("Goal=%d,%d,%d" % tuple(goal))
    return "\n".join(out) + "\n"


# ---- validator -------------------------------------------------------------------------------------------------------
@dataclass
class RoomReport:
    name: str = ""
    dims: Tuple[int, int, int] = (0, 0, 0)
    n_cells: int = 0
    n_wall: int = 0
    n_free_interior: int = 0                   # total_free_cells == max_steps (CubicEnv.py:450-459)
    findings: Dict[str, object] = field(default_factory=dict)
    errors: List[str] = field(default_factory=list)

    @property
    def ok(self) -> bool:
        return not self.errors and not self.findings


def _largest_component(free: np.ndarray) -> int:
    """Size of the largest 6-connected component of the True cells (moves are axis steps, CubicEnv.py:135-153)."""
    seen = np.zeros_like(free, dtype=bool)
    best = 0
    W, D, H = free.shape
    for sx, sy, sz in np.argwhere(free):
        if seen[sx, sy, sz]:
            continue
        n, q = 0, deque([(sx, sy, sz)])
        seen[sx, sy, sz] = True
        while q:
            x, y, z = q.popleft()
            n += 1
            for dx, dy, dz in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
                a, b, c = x + dx, y + dy, z + dz
                if 0 <= a < W and 0 <= b < D and 0 <= c < H and free[a, b, c] and not seen[a, b, c]:
                    seen[a, b, c] = True
                    q.append((a, b, c))
        best = max(best, n)
    return best


def validate_room_text(text: str, name: str = "") -> RoomReport:
    rep = RoomReport(name=name)
    # pass 1: the line-level quirks the parser accepts silently
    layers: List[int] = []
    height = None
    for raw in text.splitlines():
        line = raw.strip()
        if line.startswith("Size"):
            try:
                height = int(line.split("=")[1].split(",")[2])
            except Exception:  # noqa: BLE001
                pass
        elif line.startswith("Layer"):
            try:
                layers.append(int(line.split("=")[1]))
            except Exception:  # noqa: BLE001
                rep.errors.append(f"unparsable layer line: {line!r}")
    neg = [z for z in layers if z < 0]
    if neg and height is not None:
        rep.findings["negative_layer_index"] = {z: z % height for z in neg}
    if height is not None:
        norm = [z % height if -height <= z < height else z for z in layers]
        dup = sorted({z for z in norm if norm.count(z) > 1})
        missing = sorted(set(range(height)) - set(norm))
        if dup:
            rep.findings["duplicate_layers"] = dup
        if missing:
            rep.findings["missing_layers"] = missing          # those layers stay all-free (np.zeros, CubicEnv.py:427)
    # pass 2: the parsed grid
    try:
        room = parse_room_text(text, name=name)
    except (ValueError, IndexError) as e:
        rep.errors.append(f"{type(e).__name__}: {e}")
        return rep
    g = room.grid
    W, D, H = g.shape
    rep.dims, rep.n_cells = (W, D, H), int(g.size)
    wall = g == CUBIC_WALL
    rep.n_wall = int(wall.sum())
    stray = sorted(int(v) for v in np.unique(g) if v not in (0, CUBIC_WALL))
    if stray:
        rep.findings["stray_values"] = stray                 # kept verbatim by the loader: free cells with odd codes
    if "-2" in text.split():
        rep.findings["literal_minus_two"] = True             # a wall for CubicEnv, a FREE cell for simpleEnv (walls are +2 there)
    shell = np.ones_like(wall)
    if W > 2 and D > 2 and H > 2:
        shell[1:-1, 1:-1, 1:-1] = False
    open_cells = int((~wall & shell).sum())
    if open_cells:
        rep.findings["open_shell_cells"] = open_cells        # the agent can stand on the room edge: real OOB ray/move paths
    if W > MAX_W or D > MAX_D or H > MAX_H:
        rep.findings["exceeds_engine_limits"] = (W, D, H)
    interior_free = ~wall[1:-1, 1:-1, 1:-1] if min(W, D, H) > 2 else np.zeros((0, 0, 0), dtype=bool)
    rep.n_free_interior = int(interior_free.sum())
    if rep.n_free_interior == 0:
        rep.errors.append("no free interior cell: reset would fail in random.choice([]) (CubicEnv.py:462)")
    else:
        reach = _largest_component(~wall)
        frac = reach / max(1, int((~wall).sum()))
        if frac < 0.84:
            rep.findings["largest_connected_free_fraction"] = round(frac, 4)   # below FINISH_PERCENTAGE some starts cannot finish
    for label, pos in (("start", room.start), ("goal", room.goal)):
        if pos is not None:
            x, y, z = pos
            if not (0 <= x < W and 0 <= y < D and 0 <= z < H):
                rep.errors.append(f"{label} {pos} is outside the room")
            elif wall[x, y, z]:
                rep.findings[f"{label}_in_wall"] = pos       # the loader warns and re-picks (CubicEnv.py:464-466)
    return rep


def validate_room_file(path) -> RoomReport:
    p = Path(path)
    return validate_room_text(p.read_text(), name=p.name)


def normalise_room_text(text: str) -> str:
    """Rewrite a room file canonically (layers 0..h-1 in order, wall token 2) with an identical parsed CubicEnv grid."""
    room = parse_room_text(text)
    return room_to_text(room.grid, room.start, room.goal, cubic=True)


# ---- generator -------------------------------------------------------------------------------------------------------
def _shell(w: int, d: int, h: int) -> np.ndarray:
    g = np.zeros((w, d, h), dtype=np.int8)
    g[0], g[-1], g[:, 0], g[:, -1], g[:, :, 0], g[:, :, -1] = CUBIC_WALL, CUBIC_WALL, CUBIC_WALL, CUBIC_WALL, CUBIC_WALL, CUBIC_WALL
    return g


def generate_room(kind: str, size: Tuple[int, int, int], seed: int = 0, density: float = 0.15) -> np.ndarray:
    """``empty``: hollow box (the reference's default room, CubicEnv.py:440-448).  ``maze``: a perfect maze of one-cell
    corridors in the x-y plane (depth-first carving), full height.  ``furnished``: a hollow box with axis-aligned cuboids
    standing on the floor or hanging from the ceiling, covering about ``density`` of the floor.  All closed-shell, and the
    free space of ``empty`` and ``maze`` is connected by construction."""
    w, d, h = (int(v) for v in size)
    if min(w, d, h) < 3:
        raise ValueError("a room needs at least 3 cells per axis")
    rng = np.random.default_rng(seed)
    g = _shell(w, d, h)
    if kind == "empty":
        return g
    if kind == "maze":
        g[1:-1, 1:-1, 1:-1] = CUBIC_WALL
        cx, cy = (w - 1) // 2, (d - 1) // 2                      # corridor cells sit on odd coordinates
        seen = np.zeros((cx, cy), dtype=bool)
        stack = [(int(rng.integers(cx)), int(rng.integers(cy)))]
        seen[stack[0]] = True
        g[2 * stack[0][0] + 1, 2 * stack[0][1] + 1, 1:-1] = 0
        while stack:
            i, j = stack[-1]
            nbrs = [(i + a, j + b) for a, b in ((1, 0), (-1, 0), (0, 1), (0, -1))
                    if 0 <= i + a < cx and 0 <= j + b < cy and not seen[i + a, j + b]]
            if not nbrs:
                stack.pop()
                continue
            ni, nj = nbrs[int(rng.integers(len(nbrs)))]
            seen[ni, nj] = True
            g[i + ni + 1, j + nj + 1, 1:-1] = 0                  # the wall cell between (2i+1, 2j+1) and (2ni+1, 2nj+1)
            g[2 * ni + 1, 2 * nj + 1, 1:-1] = 0
            stack.append((ni, nj))
        return g
    if kind == "furnished":
        target = density * (w - 2) * (d - 2)
        covered, tries = 0.0, 0
        while covered < target and tries < 1000:
            tries += 1
            bw, bd = int(rng.integers(1, max(2, (w - 2) // 3))), int(rng.integers(1, max(2, (d - 2) // 3)))
            bh = int(rng.integers(1, max(2, h - 3)))
            x0, y0 = int(rng.integers(1, w - 1 - bw + 1)), int(rng.integers(1, d - 1 - bd + 1))
            if rng.random() < 0.8:
                g[x0:x0 + bw, y0:y0 + bd, 1:1 + bh] = CUBIC_WALL            # stands on the floor
            else:
                g[x0:x0 + bw, y0:y0 + bd, h - 1 - bh:h - 1] = CUBIC_WALL    # hangs from the ceiling
            covered += bw * bd
        return g
    raise ValueError(f"unknown room kind {kind!r} (empty, maze, furnished)")


# ---- furniture stamps (rooms/objects) ---------------------------------------------------------------------------------
def parse_stamp_text(text: str) -> np.ndarray:
    """A furniture stamp of ``rooms/objects`` (``Layer=k`` headers, each followed by rows of 0/2 values; no ``Size=``; the
    reference ships them as copy/paste aids that no code loads).  Returns ``stamp[x, y, k]`` with 1 where the file has a 2;
    text column -> x and text row -> y as in room files; layers the file does not mention are empty."""
    layers: Dict[int, List[List[int]]] = {}
    cur = None
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        if line.startswith("Layer"):
            cur = int(line.split("=")[1])
            layers[cur] = []
        elif cur is not None:
            layers[cur].append([int(v) for v in line.split()])
    if not layers:
        raise ValueError("stamp has no Layer= block")
    w = max(len(r) for rows in layers.values() for r in rows)
    d = max(len(rows) for rows in layers.values())
    h = max(layers) + 1
    out = np.zeros((w, d, h), dtype=np.int8)
    for k, rows in layers.items():
        if k < 0:
            raise ValueError("negative stamp layer")
        for y, row in enumerate(rows):
            for x, v in enumerate(row):
                out[x, y, k] = 1 if v == 2 else 0
    return out


def stamp_object(grid: np.ndarray, stamp: np.ndarray, x0: int, y0: int, rotate: int = 0, *, allow_overlap: bool = True) -> bool:
    """OR ``stamp`` (rotated ``rotate`` quarter turns about z) into the CubicEnv room ``grid`` with its (0, 0) corner at
    column (x0, y0); stamp layer k lands on room layer z = k (stamps start at ``Layer=1``, the first layer above the
    floor).  Returns False and leaves the room untouched when the stamp does not fit inside the shell (or, with
    ``allow_overlap=False``, would touch an existing wall cell of the interior)."""
    st = np.rot90(stamp, k=rotate % 4, axes=(0, 1))
    w, d, h = st.shape
    W, D, H = grid.shape
    if x0 < 1 or y0 < 1 or x0 + w > W - 1 or y0 + d > D - 1 or h > H - 1:
        return False
    region = grid[x0:x0 + w, y0:y0 + d, 0:h]
    if not allow_overlap and bool(((region == CUBIC_WALL) & (st == 1))[:, :, 1:].any()):
        return False
    region[st == 1] = CUBIC_WALL
    return True


def load_stamps(objects_dir) -> Dict[str, np.ndarray]:
    return {p.name: parse_stamp_text(p.read_text()) for p in sorted(Path(objects_dir).iterdir()) if p.is_file()}


def furnish_room(size: Tuple[int, int, int], stamps: Dict[str, np.ndarray], n_objects: int, seed: int = 0) -> np.ndarray:
    """A hollow box with ``n_objects`` randomly chosen, randomly rotated stamps placed without overlap (best effort)."""
    rng = np.random.default_rng(seed)
    g = _shell(*size)
    names = sorted(stamps)
    placed = tries = 0
    while placed < n_objects and tries < 200 * max(1, n_objects):
        tries += 1
        st = stamps[names[int(rng.integers(len(names)))]]
        rot = int(rng.integers(4))
        w, d = (st.shape[0], st.shape[1]) if rot % 2 == 0 else (st.shape[1], st.shape[0])
        if size[0] - 2 < w or size[1] - 2 < d:
            continue
        x0, y0 = int(rng.integers(1, size[0] - 1 - w + 1)), int(rng.integers(1, size[1] - 1 - d + 1))
        placed += stamp_object(g, st, x0, y0, rot, allow_overlap=False)
    return g


# ---- CLI -------------------------------------------------------------------------------------------------------------
def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m nav3d.room_tools", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    v = sub.add_parser("validate", help="report the quirks of room files (a file or a directory of *.txt)")
    v.add_argument("path")
    n = sub.add_parser("normalise", help="rewrite a room file canonically (same parsed grid)")
    n.add_argument("path")
    n.add_argument("-o", "--output", required=True)
    gsub = sub.add_parser("generate", help="write a procedurally generated room")
    gsub.add_argument("--kind", default="maze", choices=["empty", "maze", "furnished"])
    gsub.add_argument("--size", default="21,21,9")
    gsub.add_argument("--seed", type=int, default=0)
    gsub.add_argument("--density", type=float, default=0.15)
    gsub.add_argument("-o", "--output", required=True)
    c = sub.add_parser("compose", help="hollow box furnished with stamps from a rooms/objects directory")
    c.add_argument("--objects", required=True)
    c.add_argument("--size", default="32,32,12")
    c.add_argument("--count", type=int, default=6)
    c.add_argument("--seed", type=int, default=0)
    c.add_argument("-o", "--output", required=True)
    args = ap.parse_args(argv)
    if args.cmd == "validate":
        p = Path(args.path)
        files = sorted(p.glob("*.txt")) if p.is_dir() else [p]
        bad = 0
        for f in files:
            r = validate_room_file(f)
            status = "ok" if r.ok else ("ERROR" if r.errors else "note")
            print(f"{f.name:<48} {r.dims[0]:>3}x{r.dims[1]:<3}x{r.dims[2]:<3} free {r.n_free_interior:>6}  {status}"
                  + ("" if r.ok else f"  {r.errors or ''} {r.findings or ''}"))
            bad += bool(r.errors)
        return 1 if bad else 0
    if args.cmd == "normalise":
        Path(args.output).write_text(normalise_room_text(Path(args.path).read_text()))
        return 0
    size = tuple(int(x) for x in args.size.split(","))
    if args.cmd == "compose":
        Path(args.output).write_text(room_to_text(furnish_room(size, load_stamps(args.objects), args.count, args.seed)))
        return 0
    Path(args.output).write_text(room_to_text(generate_room(args.kind, size, args.seed, args.density)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
