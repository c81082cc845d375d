"""Room files: the reference's ``rooms/*.txt`` voxel format, parsed exactly like ``load_room``.

Replaces the parsing half of ``GridAgent.load_room`` (reference ``envs/CubicEnv.py:402-438`` and
``envs/simpleEnv.py:345-382``), which the reference re-runs on every ``reset``; here a directory is parsed once and the
dense grids are handed to ``nav3d_load_rooms`` (``include/nav3d.h``), whose CUDA kernel bit-packs them.

Format (reference ``README.md:8-17``): ``Size=w,d,h``; ``Layer z=k`` followed by ``d`` rows of ``w`` integers
(text column -> x, text row -> y); optional ``Start position=x,y,z`` / ``Goal=x,y,z``.  Quirks reproduced on purpose:

* the first matching rule wins, tested in the reference's order: blank, ``Start position``, ``Goal``, ``Size``,
  ``Layer``, data row;
* ``Layer z=-2`` is accepted and wraps like a NumPy negative index (``rooms/P3_training/kitchen2.txt:62``);
* CubicEnv rewrites 2 -> -2 and keeps every other value (so a literal ``-2`` in a file is a wall too); simpleEnv keeps
  the file's values and only ``2`` is a wall (so that file's ``-2`` cells are free there);
* a row whose length differs from ``w`` raises ``ValueError`` with the reference's message (``CubicEnv.py:435-436``);
  a row index >= d or a layer index outside [-h, h) raises ``IndexError`` like the NumPy assignment would.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

CUBIC_WALL = -2   # envs/CubicEnv.py:434
SIMPLE_WALL = 2   # envs/simpleEnv.py:281


@dataclass
class Room:
    """One parsed room: ``grid[x, y, z]`` is what the reference holds in ``self.grid`` (as int8)."""
    grid: np.ndarray
    wall_code: int
    start: Optional[Tuple[int, int, int]] = None
    goal: Optional[Tuple[int, int, int]] = None
    name: str = ""

    @property
    def dims(self) -> Tuple[int, int, int]:
        return tuple(int(v) for v in self.grid.shape)  # type: ignore[return-value]

    def free_cells(self) -> np.ndarray:
        """``possible_start_pose`` (``CubicEnv.py:450-457``): interior non-wall cells, x-major then y then z, int32 [n,3]."""
        g = self.grid
        inner = g[1:-1, 1:-1, 1:-1] != self.wall_code
        idx = np.argwhere(inner)            # C order == x outer, y, z inner
        return (idx + 1).astype(np.int32)


def parse_room_text(text: str, *, simple: bool = False, name: str = "") -> Room:
    grid = None
    width = depth = height = 0
    start = goal = None
    z_index = None
    row_index = 0
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        if line.startswith("Start position"):
            sx, sy, sz = map(int, line.split("=")[1].split(","))
            start = (sx, sy, sz)
        elif line.startswith("Goal"):
            gx, gy, gz = map(int, line.split("=")[1].split(","))
            goal = (gx, gy, gz)
        elif line.startswith("Size"):
            dims = line.split("=")[1].split(",")
            width, depth, height = int(dims[0]), int(dims[1]), int(dims[2])
            grid = np.zeros((width, depth, height), dtype=np.int64)
        elif line.startswith("Layer"):
            z_index = int(line.split("=")[1])
            row_index = 0
        else:
            values = list(map(int, line.split()))
            if not simple:
                values = [v if v != 2 else -2 for v in values]
            if grid is None:
                raise ValueError(f"Line '{line}' appears before the Size= line.")
            if len(values) != width:
                raise ValueError(f"Line '{line}' has {len(values)} values, but width is {width} for layer {z_index}, row {row_index}.")
            if z_index is None:
                raise ValueError(f"Line '{line}' appears before any Layer line.")
            grid[:, row_index, z_index] = values      # IndexError for bad row / layer, negative layer wraps
            row_index += 1
    if grid is None:
        raise ValueError("room has no Size= line")
    if grid.min() < -128 or grid.max() > 127:
        raise ValueError("room cell values must fit int8")
    return Room(grid=np.ascontiguousarray(grid.astype(np.int8)), wall_code=SIMPLE_WALL if simple else CUBIC_WALL,
                start=start, goal=goal, name=name)


def load_room_file(path, *, simple: bool = False) -> Room:
    p = Path(path)
    with open(p, "r") as f:
        return parse_room_text(f.read(), simple=simple, name=p.name)


def list_room_files(room_path) -> List[Path]:
    """``list(Path(room_path).glob('*.txt'))`` — directory-iteration order, NOT sorted, exactly as ``CubicEnv.py:64-66``.

    The order matters: ``random.choice(self.rooms)`` indexes this list (``:407``)."""
    return list(Path(room_path).glob("*.txt"))


def load_room_dir(room_path, *, simple: bool = False, sort: bool = False) -> List[Room]:
    files = list_room_files(room_path)
    if sort:
        files = sorted(files)
    if not files:
        raise FileNotFoundError(f"no *.txt rooms under {room_path}")
    return [load_room_file(p, simple=simple) for p in files]


def default_box_room(width: int = 20, depth: int = 20, height: int = 12, *, simple: bool = False) -> Room:
    """The hollow box the reference builds when ``room_path`` is None (``CubicEnv.py:440-448``)."""
    wall = SIMPLE_WALL if simple else CUBIC_WALL
    g = np.zeros((width, depth, height), dtype=np.int8)
    g[0, :, :] = wall; g[-1, :, :] = wall
    g[:, 0, :] = wall; g[:, -1, :] = wall
    g[:, :, 0] = wall; g[:, :, -1] = wall
    return Room(grid=g, wall_code=wall, name="<box>")


def rooms_from_grids(grids: Sequence[np.ndarray], wall_code: int = CUBIC_WALL) -> List[Room]:
    return [Room(grid=np.ascontiguousarray(np.asarray(g, dtype=np.int8)), wall_code=wall_code, name=f"<grid{i}>")
            for i, g in enumerate(grids)]
