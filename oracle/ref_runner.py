"""Runs the UNMODIFIED reference env (``oracle/_ref/envs/CubicEnv.py``) as a CPU baseline.  TEST/BENCH INFRASTRUCTURE ONLY.

``oracle/_ref/envs/`` holds byte-identical copies of the reference's ``envs/CubicEnv.py`` and ``envs/simpleEnv.py``, made by
``__graft_entry__.build()`` in the container that has ``/root/reference`` (git-ignored, shipped to the GPU box like the
built ``.so`` files; never committed).  The two files import ``gymnasium`` and ``matplotlib``, which this image does not
have: the stand-ins below are SURVEY.md Appendix A's ``sys.modules`` shims — the env only needs ``gym.Env.reset(seed=)``,
``spaces.Discrete(n).n`` and an importable ``matplotlib.pyplot``.

Workload = BASELINE.md §3 (BASELINE.json configs[0]): ``GridAgent(room_path=rooms/P1_training, local_map_length=10)``,
``reset(seed=42 + rank)``, uniform random actions from ``np.random.default_rng(rank)``, ``reset()`` whenever the episode
ends (so the per-reset file parse and free-cell scan of ``load_room`` :402-473 are inside the timed region), stdout
redirected (the env prints at every episode end, :217, :222).  One env per worker process, like the reference's own
``SubprocVecEnv`` (``train/Grid_Train.py:191-192``).

Nothing here imports the product package (``nav3d``) or its library.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"


def available() -> bool:
    return (REF_DIR / "envs" / "CubicEnv.py").exists()


def install_shims() -> None:
    """Stand-ins for the absent third-party imports of the reference env files."""
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:                                   # gym.Env.reset only seeds self.np_random, which the env never reads
            def reset(self, *, seed=None, options=None):
                pass

        class Discrete:
            def __init__(self, n):
                self.n = n                           # read at CubicEnv.py:284

        class Box:
            def __init__(self, low, high, dtype=None, shape=None):
                self.low, self.high, self.dtype = low, high, dtype
                self.shape = getattr(low, "shape", shape)

        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Discrete, spaces.Box = Discrete, Box
        gym.Env, gym.spaces = Env, spaces
        sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces})
    for m in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if not hasattr(sys.modules["mpl_toolkits.mplot3d"], "Axes3D"):
        sys.modules["mpl_toolkits.mplot3d"].Axes3D = object


def load_grid_agent(simple: bool = False):
    """The reference's ``GridAgent`` class, from the unmodified copy under oracle/_ref."""
    install_shims()
    import importlib.util
    name = "simpleEnv" if simple else "CubicEnv"
    spec = importlib.util.spec_from_file_location(f"nav3d_ref_envs_{name}", REF_DIR / "envs" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.GridAgent


def worker_rollout(args):
    """(room_dir, L, n_steps, rank) -> (steps done, seconds).  Import and construction are outside the timer; the resets
    (incl. the first) are inside, as BASELINE.md §3 says."""
    room_dir, L, n_steps, rank = args
    import numpy as np
    GridAgent = load_grid_agent()
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        env = GridAgent(room_path=str(room_dir), local_map_length=L)
        rng = np.random.default_rng(rank)
        actions = rng.integers(0, 6, size=n_steps)
        t0 = time.perf_counter()
        env.reset(seed=42 + rank)
        for a in actions:
            _, _, term, trunc, _ = env.step(int(a))
            if term or trunc:
                env.reset()
        dt = time.perf_counter() - t0
    return n_steps, dt


def reference_rate(room_dir, L, steps_per_worker, workers):
    """`workers` processes, one reference env each; returns (env-steps/s summed over the workers, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(worker_rollout, [(str(room_dir), L, steps_per_worker, i) for i in range(workers)])
    wall = time.perf_counter() - t0
    return sum(n / t for n, t in res), wall


if __name__ == "__main__":
    root = HERE.parent
    w = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    rate, wall = reference_rate(root / "rooms" / "P1_training", 10, n, w)
    print(f"reference CubicEnv: {w} processes x {n} steps: {rate:.0f} env-steps/s ({wall:.1f} s)")
