"""``GridAgent`` of the reference's older env variant ``envs/simpleEnv.py`` (:13-520) as a scalar facade over the CUDA
engine (``NAV3D_ENV_SIMPLE``).  Differences from the reference that are deliberate: ``reset`` returns ``(obs, {})``
instead of ``None`` (the reference's ``reset`` forgets its return statement, ``simpleEnv.py:79-107``; ``obs`` is what a
``get_obs()`` call right after the reference's reset returns)."""
from __future__ import annotations

import random
from pathlib import Path

import numpy as np
import torch

from nav3d import _lib
from nav3d.engine import Engine
from nav3d.rooms import default_box_room, load_room_file
from nav3d.spaces import simple_spaces

NN_SIZES = [[32, 32], [64, 64], [128, 128], [256, 256]]
FINISH_PERCENTAGE = 0.8      # defined but unused by the reference (simpleEnv.py:10)
SPOT_GOAL_HEIGTH = 5         # simpleEnv.py:11


class GridAgent:
    def __init__(self, grid=None, max_steps=2000, width: int = 20, depth: int = 20, height: int = 12,
                 cell_size: float = 0.25, local_map_length=4, room_path=None, render_mode: str = None, device: int = 0):
        self.width, self.depth, self.height = width, depth, height
        self.cell_size, self.local_map_length, self.max_steps = cell_size, local_map_length, max_steps
        self.rooms = None
        self.valid_facings = {0: "north", 1: "east", 2: "south", 3: "west"}
        self.action_space, self.observation_space = simple_spaces(local_map_length)
        if room_path is not None:
            self.rooms = list(Path(room_path).glob("*.txt"))
        self.render_mode = render_mode
        self._device, self._engine, self._room_key = device, None, None

    def _ensure_engine(self):
        key = None if self.rooms is None else tuple(str(p) for p in self.rooms)
        if self._engine is not None and key == self._room_key:
            return
        parsed = ([default_box_room(self.width, self.depth, self.height, simple=True)] if self.rooms is None
                  else [load_room_file(p, simple=True) for p in self.rooms])
        self._engine = Engine(1, parsed, local_map_length=int(self.local_map_length), auto_reset=False,
                              device=self._device, env_kind=_lib.ENV_SIMPLE, cell_size=float(self.cell_size))
        self._loaded_rooms, self._room_key = parsed, key
        dev, d = self._engine.device, self._engine.obs_dim
        self._obs = torch.zeros((1, d), dtype=torch.float32, device=dev)
        self._rew = torch.zeros(1, dtype=torch.float32, device=dev)
        self._rew64 = torch.zeros(1, dtype=torch.float64, device=dev)
        self._term = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._act = torch.zeros(1, dtype=torch.int64, device=dev)

    def _sync_attrs(self):
        s = self._engine.get_state()[0].cpu().numpy()
        self.x, self.y, self.z, self.facing = int(s[0]), int(s[1]), int(s[2]), int(s[3])
        self.visited_count, self.bump_count, self.step_count = int(s[4]), int(s[5]), int(s[6])
        self.gx, self.gy, self.gz = int(s[7]), int(s[8]), int(s[9])
        self.done, self.last_action = bool(s[10]), int(s[12])

    def reset(self, seed: int | None = None, options: dict | None = None):
        # like the reference, reset does NOT seed `random` (simpleEnv.py:79-81); it consumes the global stream
        self._ensure_engine()
        rooms = self._loaded_rooms
        ri = random.choice(range(len(rooms))) if self.rooms is not None else 0          # :351
        room = rooms[ri]
        cells = room.free_cells()
        self.width, self.depth, self.height = room.dims
        self.grid = room.grid.astype(int)
        self.total_free_cells = self.max_steps = len(cells)                                # :395-404

        def index_of(cell):
            hit = np.nonzero((cells == np.asarray(cell)).all(axis=1))[0]
            if len(hit) == 0:
                raise NotImplementedError("a file-provided start/goal on the boundary shell is not supported by the engine")
            return int(hit[0])

        if room.start is None:
            k = random.choice(range(len(cells)))                                           # :406
        elif self.grid[room.start] == 2:                                                   # :409-412
            print(f"Warning: Provided start position ({room.start[0]},{room.start[1]},{room.start[2]}) is a wall. Choosing a random valid start position.")
            k = random.choice(range(len(cells)))
        else:
            k = index_of(room.start)
        if room.goal is None:
            kg = random.choice(range(len(cells)))                                          # :417
        elif self.grid[room.goal] == 2:                                                    # :420-423
            print("Warning: Provided GOAL position is a wall. Choosing a random valid start position.")
            kg = random.choice(range(len(cells)))
        else:
            kg = index_of(room.goal)
        self._engine.reset(self._obs, picks=torch.tensor([[ri, k, kg]], dtype=torch.int32))
        self._sync_attrs()
        return self._obs[0].cpu().numpy(), {}

    def step(self, action: int):
        self._act[0] = int(action)
        self._engine.step(self._act, self._obs, self._rew, self._term, self._trunc, reward64=self._rew64)
        obs = self._obs[0].cpu().numpy()
        reward = float(self._rew64[0].item())
        terminated, truncated = bool(self._term[0].item()), bool(self._trunc[0].item())
        self._sync_attrs()
        if reward >= 99.0 - 10.2:                                                          # goal branch fired this step (:206-207)
            print(f"Finished after: {self.step_count} Steps : {self.bump_count} bumps")
        if truncated:                                                                      # :212
            print(f"Truncated after: {self.step_count} Steps, with: {self.bump_count} Bumps and {self.visited_count} cells discovered")
        return obs, reward, terminated, truncated, {}

    def get_obs(self):
        return self._obs[0].cpu().numpy()

    @property
    def internal_grid(self) -> np.ndarray:
        return self._engine.get_grid(0).astype(int)

    def get_position(self):
        return (self.x, self.y, self.z)

    def render(self):
        if self.render_mode == "human":
            print(f"--- Step: {self.step_count}, Pos: ({self.x}, {self.y}, {self.z}), Facing: {self.valid_facings[self.facing]} ---")

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
