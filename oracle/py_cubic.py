"""Pure Python/NumPy restatement of the reference's CubicEnv ``GridAgent``.  TEST INFRASTRUCTURE ONLY.

Why a second oracle next to ``nav3d_oracle.c``: the reference is an interpreted Python class stepped one env per OS
process (``train/Grid_Train.py:118-126, :191-192``).  ``/root/reference`` does not exist on the GPU box, so the
*reference arm* of ``bench.py`` (``--impl reference``) and its ``cpu_baseline`` time THIS port — same language, same
per-cell Python loops, same NumPy calls per step — as the stand-in for "the reference Python env on the box's host cores".
The C oracle is the fast checker; this file is the faithful-cost one.  Both are pinned to the reference by
``tests/test_oracle_golden.py`` / ``tests/test_py_oracle.py`` against ``tests/golden/cubic_traces.npz``.

Follows ``envs/CubicEnv.py``: reset :77-108, step :110-132, do_action :134-166, compute_reward :169-224,
_get_3d_local_map :229-251, get_obs :254-312, _mark_visited :322-343, _sense_direction :345-397.
The (room, start) picks are injected (the reference draws them from CPython's ``random``: :407, :462).
"""
from __future__ import annotations

import numpy as np

FINISH_FRACTION = 0.84                      # CubicEnv.py:12
_HEADINGS = ((0, 1), (1, 0), (0, -1), (-1, 0))   # facing 0..3 = N(+y), E(+x), S(-y), W(-x)


class PyCubic:
    def __init__(self, local_map_length=4, crash_penalty=-2.0):
        self.L = local_map_length
        self.crash_penalty = crash_penalty
        self.total_free = 1

    # -- reset ---------------------------------------------------------------------------------------------------
    def reset(self, grid, total_free, start):
        """``grid``: int array [W, D, H] with -2 walls (what load_room leaves in self.grid)."""
        self.g = grid
        self.W, self.D, self.H = grid.shape
        self.total_free = total_free
        self.max_steps = total_free                      # :459
        self.x, self.y, self.z = (int(v) for v in start)
        self.ig = np.full(grid.shape, -1, dtype=int)     # :84
        self.ig[self.x, self.y, self.z] = 1              # :85
        self.visited = 1
        self.steps = 0
        self.bumps = 0
        self.facing = 0
        self.last_action = 0
        self.done = self.explored = self.bumped = self.last_bump = False
        self.near_wall = self.was_near_wall = False
        self.down = 0
        return self._observe()

    # -- step ----------------------------------------------------------------------------------------------------
    def step(self, action):
        if self.near_wall:                               # :111-113
            self.was_near_wall = True
            self.near_wall = False
        self.steps += 1
        truncated = self.steps >= self.max_steps         # :116
        self._move(int(action))
        obs = self._observe()
        reward = self._reward(int(action), truncated)
        self.last_action = int(action)                   # :124
        return obs, reward, self.done, truncated

    def _inside(self, x, y, z):
        return 0 <= x < self.W and 0 <= y < self.D and 0 <= z < self.H

    def _move(self, a):                                  # do_action :134-166 with _mark_visited :322-343 inlined
        vx = vy = vz = 0
        if a < 4:
            self.facing = (self.facing + a) % 4          # the four tables of :135-140 are rotations of one another
            vx, vy = _HEADINGS[self.facing]
        else:
            vz = 1 if a == 4 else -1
        tx, ty, tz = self.x + vx, self.y + vy, self.z + vz
        ok = self._inside(tx, ty, tz) and self.g[tx, ty, tz] != -2
        if ok:
            v = self.ig[tx, ty, tz]
            if v == 0:
                self.ig[tx, ty, tz] = 1
                self.visited += 1
                self.explored = True
            elif v > 0:
                self.ig[tx, ty, tz] = v + 1
            self.x, self.y, self.z = tx, ty, tz
        else:
            self.bumped = True
        if self._inside(self.x, self.y, self.z) and self.g[self.x, self.y, self.z] != -2:
            self.ig[self.x, self.y, self.z] += 1         # :165-166

    def _ray(self, dx, dy, dz):                          # _sense_direction :345-397
        free = 0
        hit = False
        for s in range(1, self.L + 1):
            nx, ny, nz = self.x + dx * s, self.y + dy * s, self.z + dz * s
            if not self._inside(nx, ny, nz):
                break
            if not hit:
                if self.g[nx, ny, nz] == -2:
                    self.ig[nx, ny, nz] = -2
                    hit = True
                    if s == 1:
                        self.near_wall = True
                else:
                    free += 1
                    if self.ig[nx, ny, nz] == -1:
                        self.ig[nx, ny, nz] = 0
        if dz == -1:
            self.down = free

    def _observe(self):                                  # get_obs :254-312
        fx, fy = _HEADINGS[self.facing]
        lx, ly = _HEADINGS[(self.facing + 3) % 4]
        for d in ((fx, fy, 0), (lx, ly, 0), (-lx, -ly, 0), (-fx, -fy, 0), (0, 0, 1), (0, 0, -1)):   # :264
            self._ray(*d)
        win = np.full((4, 4, 4), -1, dtype=np.float32)   # _get_3d_local_map :229-251
        for i in range(4):
            gx = self.x + i - 2
            for j in range(4):
                gy = self.y + j - 2
                for k in range(4):
                    gz = self.z + k - 2
                    if self._inside(gx, gy, gz):
                        win[i, j, k] = self.ig[gx, gy, gz]
        flat = np.clip(win.flatten(), -2, 20.0)          # :273-275
        flat = (flat + 2) / (20.0 + 2)
        heading = np.zeros(4, dtype=np.float32)
        heading[self.facing] = 1.0
        scal = np.array([float(self.last_action) / 5, float(self.was_near_wall), float(self.last_bump),
                         float(self.down) / self.L], dtype=np.float32)
        frac = np.array([self.visited / self.total_free], dtype=np.float32)
        obs = np.concatenate([flat, heading, scal, frac])
        return np.pad(obs, (0, 80 - obs.shape[0]), "constant", constant_values=0)

    def _reward(self, a, truncated):                     # compute_reward :169-224
        r = -0.05
        r -= min(self.ig[self.x, self.y, self.z] * 0.02, 0.5)
        if self.bumped:
            self.bumped = False
            self.last_bump = True
            self.bumps += 1
            r += self.crash_penalty
        else:
            self.last_bump = False
            if self.was_near_wall:
                self.was_near_wall = False
                r += 0.15
            if self.last_action != 2 and a == self.last_action and self.last_action < 4:
                r += 0.05
            if self.last_action == 2 and a == 2:
                r -= 0.5
        if self.explored:
            self.explored = False
            r += 1.0
        if self.visited / self.total_free >= FINISH_FRACTION:
            self.done = True
            r += 100.0
        if truncated:
            r += -5.0
        return r


def free_cells(grid):
    """possible_start_pose (:450-457) as an int array [n, 3]."""
    return np.argwhere(grid[1:-1, 1:-1, 1:-1] != -2) + 1


def worker_rollout(args):
    """One reference-style worker (one env per process, ``SubprocVecEnv``): ``n_steps`` random-action steps with reset on
    done, over the given room grids.  Returns (env steps, seconds)."""
    import time
    grids, L, n_steps, seed = args
    rng = np.random.default_rng(seed)
    cells = [free_cells(g) for g in grids]
    env = PyCubic(L)

    def do_reset():
        r = int(rng.integers(0, len(grids)))
        env.reset(grids[r], len(cells[r]), cells[r][int(rng.integers(0, len(cells[r])))])

    do_reset()
    actions = rng.integers(0, 6, size=n_steps)
    t0 = time.perf_counter()
    for a in actions:
        _, _, term, trunc = env.step(a)
        if term or trunc:
            do_reset()
    return n_steps, time.perf_counter() - t0
