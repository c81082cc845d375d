"""``GridAgent`` — scalar gym-style facade with the constructor, methods and public attributes of the reference's
``envs/CubicEnv.py::GridAgent`` (:15-538), executed by the batched CUDA engine with a batch of one.

It exists so that the reference's own scripts (``train/Grid_Train.py:94-126``, ``train/evaluate_grid.py:28-72``) keep
working unchanged; throughput comes from ``nav3d.BatchedCubicEnv``, not from this class (one kernel launch per step)."""
from __future__ import annotations

import random
from pathlib import Path

import numpy as np
import torch

from nav3d.engine import Engine
from nav3d.rooms import default_box_room, load_room_file
from nav3d.spaces import cubic_spaces

try:  # gymnasium is optional (absent in the build image)
    import gymnasium as _gym
    _Base = _gym.Env
except Exception:  # noqa: BLE001
    _Base = object

NN_SIZES = [[32, 32], [64, 64], [128, 128], [256, 256]]   # CubicEnv.py:11
FINISH_PERCENTAGE = 0.84                                    # CubicEnv.py:12


class GridAgent(_Base):
    def __init__(self, grid=None, max_steps=2000, width: int = 20, depth: int = 20, height: int = 12,
                 cell_size: float = 0.25, local_map_length=4, room_path=None, render_mode: str = None,
                 crash_penalty: float = -2.0, device: int = 0):
        if _Base is not object:
            super().__init__()
        self.width, self.depth, self.height = width, depth, height
        self.cell_size = cell_size
        self.local_map_length = local_map_length
        self.max_steps = max_steps
        self.rooms = None
        self.valid_facings = {0: "north", 1: "east", 2: "south", 3: "west"}
        self.crash_penalty = crash_penalty
        self.action_map = {0: [0], 1: [+1], 2: [+2], 3: [+3], 4: [0, 0, 1], 5: [0, 0, -1]}
        self.action_space, self.observation_space = cubic_spaces()
        if room_path is not None:
            self.rooms = list(Path(room_path).glob("*.txt"))          # directory order, as CubicEnv.py:64-66
        self.render_mode = render_mode
        self.total_free_cells = 1                                        # CubicEnv.py:74
        self._device = device
        self._engine = None
        self._loaded_rooms = None
        self._room_key = None

    # ---- engine plumbing ---------------------------------------------------------------------------------------
    def _ensure_engine(self):
        key = None if self.rooms is None else tuple(str(p) for p in self.rooms)
        if self._engine is not None and key == self._room_key:
            return
        if self.rooms is None:
            parsed = [default_box_room(self.width, self.depth, self.height)]     # CubicEnv.py:440-448
        else:
            if not self.rooms:
                raise IndexError("Cannot choose from an empty sequence")        # random.choice([]) in the reference
            parsed = [load_room_file(p) for p in self.rooms]                      # ValueError on bad rows (:435-436)
        if self._engine is not None:
            self._engine.close()
        self._engine = Engine(1, parsed, local_map_length=int(self.local_map_length),
                              crash_penalty=float(self.crash_penalty), auto_reset=False, device=self._device)
        self._loaded_rooms, self._room_key = parsed, key
        dev = self._engine.device
        self._obs = torch.zeros((1, 80), dtype=torch.float32, device=dev)
        self._rew = torch.zeros(1, dtype=torch.float32, device=dev)
        self._rew64 = torch.zeros(1, dtype=torch.float64, device=dev)
        self._term = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._act = torch.zeros(1, dtype=torch.int64, device=dev)

    def _sync_attrs(self):
        s = self._engine.get_state()[0].cpu().numpy()
        self.x, self.y, self.z, self.facing = int(s[0]), int(s[1]), int(s[2]), int(s[3])
        self.visited_count, self.bump_count, self.step_count = int(s[4]), int(s[5]), int(s[6])
        self.near_wall, self.was_near_wall, self.last_bump = bool(s[7]), bool(s[8]), bool(s[9])
        self.done = bool(s[10])
        self.cells_insight_down, self.last_action = int(s[11]), int(s[12])
        self._room_index = int(s[13])

    # ---- gym API -----------------------------------------------------------------------------------------------
    def reset(self, seed: int | None = None, options: dict | None = None):
        if _Base is not object:
            super().reset(seed=seed)
        random.seed(seed)                     # CubicEnv.py:79-80: the GLOBAL generators, on purpose
        np.random.seed(seed)
        self._ensure_engine()
        rooms = self._loaded_rooms
        ri = random.choice(range(len(rooms))) if self.rooms is not None else 0       # random.choice(self.rooms) (:407)
        room = rooms[ri]
        cells = room.free_cells()
        self.width, self.depth, self.height = room.dims
        self.grid = room.grid.astype(int)
        self.total_free_cells = len(cells)
        self.max_steps = self.total_free_cells                                         # :459
        start = room.start
        if start is None:
            k = random.choice(range(len(cells)))                                       # :462
        else:
            sx, sy, sz = start
            if self.grid[sx][sy][sz] == -2:                                            # :464-466
                print(f"Warning: Provided start position ({sx},{sy},{sz}) is a wall. Choosing a random valid start position.")
                k = random.choice(range(len(cells)))
            else:
                hit = np.nonzero((cells == np.asarray(start)).all(axis=1))[0]
                if len(hit) == 0:
                    raise NotImplementedError("a file-provided start on the room's boundary shell is not supported by the engine")
                k = int(hit[0])
        self.gx, self.gy, self.gz = room.goal if room.goal is not None else (0, 0, 0)  # :468-471 (unused by CubicEnv)
        picks = torch.tensor([[ri, k]], dtype=torch.int32)
        self._engine.reset(self._obs, picks=picks)
        self._sync_attrs()
        self.explored = self.bumped = False
        return self._obs[0].cpu().numpy(), {}

    def step(self, action: int):
        if self._engine is None:
            raise AttributeError("'GridAgent' object has no attribute 'internal_grid' (call reset() first)")
        self._act[0] = int(action)
        self._engine.step(self._act, self._obs, self._rew, self._term, self._trunc, reward64=self._rew64)
        obs = self._obs[0].cpu().numpy()
        reward = float(self._rew64[0].item())
        terminated, truncated = bool(self._term[0].item()), bool(self._trunc[0].item())
        self._sync_attrs()
        pct = self.visited_count / self.total_free_cells
        if terminated:                                                                  # messages of :217 and :222
            print(f"Goal Reached! Explored {pct*100:.2f}% after {self.step_count} Steps.")
        if truncated:
            print(f"Truncated after: {self.step_count} Steps, with: {self.bump_count} Bumps and {pct*100:.2f}% of cells discovered")
        if self.render_mode == "human":
            self.render()
        return obs, reward, terminated, truncated, {}

    def get_obs(self):
        """The current observation (the reference's get_obs re-senses, which changes nothing right after a step)."""
        return self._obs[0].cpu().numpy()

    @property
    def internal_grid(self) -> np.ndarray:
        return self._engine.get_grid(0).astype(int)

    def get_position(self):
        return (self.x, self.y, self.z)

    def _is_blocked(self, x: int, y: int, z: int) -> bool:
        return bool(self.internal_grid[x][y][z] == -2)

    # ---- rendering (CubicEnv.py:475-538) ---------------------------------------------------------------------------
    def render(self):
        if self.render_mode == "human":
            self._render_text()
        elif self.render_mode == "matplotlib":
            self._render_matplotlib()

    def _render_text(self):
        print(f"--- Step: {self.step_count}, Pos: ({self.x}, {self.y}, {self.z}), Facing: {self.valid_facings[self.facing]} ---")
        layer = np.copy(self.internal_grid[:, :, self.z])
        layer[self.x, self.y] = 9
        glyph = {9: "A ", -2: "# ", 0: "o "}
        for yy in range(self.depth):
            print("".join(glyph.get(int(v), ". " if v >= 1 else "? ") for v in layer[:, yy]))
        print("-" * (self.width * 2))

    def _render_matplotlib(self):
        import matplotlib.pyplot as plt  # optional dependency, imported on use
        if getattr(self, "fig", None) is None:
            plt.ion()
            self.fig = plt.figure(figsize=(10, 8))
            self.ax = self.fig.add_subplot(111, projection="3d")
        ax = self.ax
        ax.clear()
        ax.set_title(f"3D Grid Exploration (Step: {self.step_count})")
        ax.set_xlim(-1, self.width); ax.set_ylim(-1, self.depth); ax.set_zlim(-1, self.height)
        wx, wy, wz = np.where(self.internal_grid[:, :, 1:] == -2)
        ax.scatter(wx, wy, wz + 1, c="black", marker="s", s=100, label="Walls (Known)")
        ax.scatter(self.x, self.y, self.z, c="red", marker="^", s=200, label="Agent")
        dx, dy = [(0, 1), (1, 0), (0, -1), (-1, 0)][self.facing]
        ax.quiver(self.x, self.y, self.z, dx, dy, 0, length=1, color="red", linewidth=3, arrow_length_ratio=0.3)
        plt.legend(); plt.draw(); plt.pause(0.01)

    def close(self):
        if getattr(self, "fig", None) is not None:
            import matplotlib.pyplot as plt
            plt.close(self.fig)
            self.fig = self.ax = None
        if self._engine is not None:
            self._engine.close()
            self._engine = None
