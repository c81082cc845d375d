#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference env.

Runs only in the build container (needs /root/reference, which does not travel to the GPU box).
The reference imports gymnasium and matplotlib, neither of which is installed; they are replaced by
inert stand-ins in sys.modules before the reference files are imported (the env only needs
gym.Env.reset(seed=...), spaces.Discrete(n).n and spaces.Box).

    python tests/golden/make_golden.py

writes
    tests/golden/rooms.json            per-room known answers (dims, interior-free count, grid hashes)
    tests/golden/cubic_traces.npz      CubicEnv lock-step traces (state, reward, obs, final knowledge grid)
    tests/golden/simple_traces.npz     simpleEnv lock-step traces
    tests/golden/seeded_picks.json     (room, start) picks of reset(seed=s) — pins the MT19937 emulation

Rooms are addressed relative to the repo's rooms/ directory (a verbatim data copy of the reference's
rooms/P{1,2,3}_{training,evaluate}); the generator reads them from /root/reference/rooms to be sure the
fixtures describe the reference's own files.
"""
import contextlib
import hashlib
import io
import json
import random as pyrandom
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
ROOM_DIRS = ["P1_training", "P1_evaluate", "P2_training", "P2_evaluate", "P3_training", "P3_evaluate"]


def install_shims():
    gym = types.ModuleType("gymnasium")

    class Env:
        def reset(self, *, seed=None, options=None):
            pass

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Box:
        def __init__(self, low, high, dtype=None, shape=None):
            self.low, self.high, self.dtype = low, high, dtype
            self.shape = np.shape(low)

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Discrete, spaces.Box = Discrete, Box
    gym.Env, gym.spaces = Env, spaces
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces})
    for m in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules[m] = types.ModuleType(m)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.path.insert(0, str(REF))


def sha16(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:16]


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# --------------------------------------------------------------------------------------------
def room_answers(CubicAgent, SimpleAgent):
    out = {}
    for d in ROOM_DIRS:
        for p in sorted((REF / "rooms" / d).glob("*.txt")):
            key = f"{d}/{p.name}"
            env = CubicAgent(room_path=str(p.parent), local_map_length=4)
            env.rooms = [p]
            with quiet():
                env.reset(seed=0)
            g = env.grid
            shell_closed = bool(
                (g[0] == -2).all() and (g[-1] == -2).all() and (g[:, 0] == -2).all()
                and (g[:, -1] == -2).all() and (g[:, :, 0] == -2).all() and (g[:, :, -1] == -2).all())
            senv = SimpleAgent(room_path=str(p.parent), local_map_length=4)
            senv.rooms = [p]
            pyrandom.seed(0)
            with quiet():
                senv.reset()
            out[key] = {
                "dims": [int(env.width), int(env.depth), int(env.height)],
                "total_free_cells": int(env.total_free_cells),
                "n_wall_cells": int((g == -2).sum()),
                "cubic_grid_sha": sha16(np.ascontiguousarray(g.astype(np.int8)).tobytes()),
                "shell_closed": shell_closed,
                "simple_total_free_cells": int(senv.total_free_cells),
                "simple_grid_sha": sha16(np.ascontiguousarray(senv.grid.astype(np.int8)).tobytes()),
            }
    return out


def cubic_state_row(env, terminated, truncated):
    return [env.x, env.y, env.z, env.facing, env.visited_count, env.bump_count, env.step_count,
            int(terminated), int(truncated)]


def greedy_action(env, rng):
    """A coverage policy (NOT part of the reference): breadth-first search over the true grid to the nearest
    free cell that has not been visited yet and take the first move of that path; 3 % random moves.
    Only used to produce traces that reach the 84 % termination branch."""
    from collections import deque
    table = {0: [(0, 1, 0), (1, 0, 0), (0, -1, 0), (-1, 0, 0)],
             1: [(1, 0, 0), (0, -1, 0), (-1, 0, 0), (0, 1, 0)],
             2: [(0, -1, 0), (-1, 0, 0), (0, 1, 0), (1, 0, 0)],
             3: [(-1, 0, 0), (0, 1, 0), (1, 0, 0), (0, -1, 0)]}
    if rng.random() < 0.03:
        return int(rng.integers(0, 6))
    W, D, H = env.width, env.depth, env.height
    src = (env.x, env.y, env.z)
    first = {src: None}
    dq = deque([src])
    moves = [(0, 1, 0), (1, 0, 0), (0, -1, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1)]
    goal_move = None
    while dq:
        c = dq.popleft()
        if c != src and env.internal_grid[c] <= 0:
            goal_move = first[c]
            break
        for v in moves:
            t = (c[0] + v[0], c[1] + v[1], c[2] + v[2])
            if not (0 <= t[0] < W and 0 <= t[1] < D and 0 <= t[2] < H):
                continue
            if env.grid[t] == -2 or t in first:
                continue
            first[t] = v if c == src else first[c]
            dq.append(t)
    if goal_move is None:
        return int(rng.integers(0, 6))
    if goal_move[2] != 0:
        return 4 if goal_move[2] > 0 else 5
    for a in range(4):
        if table[a][env.facing] == goal_move:
            return a
    raise AssertionError


def cubic_trace(CubicAgent, room, L, seed, aseed, n, policy="random", crash_penalty=-2.0):
    p = REF / "rooms" / room
    env = CubicAgent(room_path=str(p.parent), local_map_length=L, crash_penalty=crash_penalty)
    env.rooms = [p]
    with quiet():
        obs0, _ = env.reset(seed=seed)
    start = (env.x, env.y, env.z)
    rng = np.random.default_rng(aseed)
    if policy == "random":
        actions = rng.integers(0, 6, size=n)
    obs = np.zeros((n + 1, 80), np.float32)
    obs[0] = obs0
    state = np.zeros((n, 9), np.int32)
    rew = np.zeros(n, np.float64)
    acts = np.zeros(n, np.int64)
    flags = np.zeros((n, 3), np.int8)  # near_wall, was_near_wall, last_bump after the step
    for t in range(n):
        a = int(actions[t]) if policy == "random" else greedy_action(env, rng)
        acts[t] = a
        with quiet():
            o, r, term, trunc, _ = env.step(a)
        obs[t + 1] = o
        rew[t] = r
        state[t] = cubic_state_row(env, term, trunc)
        flags[t] = [env.near_wall, env.was_near_wall, env.last_bump]
    ig = env.internal_grid.astype(np.int32)
    return {
        "room": room, "L": L, "seed": seed, "aseed": aseed, "n": n, "policy": policy,
        "crash_penalty": crash_penalty,
        "start": np.array(start, np.int32), "actions": acts, "state": state, "reward": rew, "obs": obs,
        "flags": flags, "final_ig": ig, "total_free": int(env.total_free_cells),
        "obs_sha": sha16(obs.tobytes()), "state_sha": sha16(state.tobytes()),
        "reward_sha": sha16(rew.astype(np.float32).tobytes()),
    }


def simple_trace(SimpleAgent, room, L, seed, aseed, n, goal_bias=0.0):
    p = REF / "rooms" / room
    env = SimpleAgent(room_path=str(p.parent), local_map_length=L)
    env.rooms = [p]
    pyrandom.seed(seed)          # simpleEnv.reset does not seed `random` itself (simpleEnv.py:79-107)
    with quiet():
        env.reset()
    start = (env.x, env.y, env.z)
    goal = (env.gx, env.gy, env.gz)
    rng = np.random.default_rng(aseed)
    obs = np.zeros((n + 1, 6 * L + 7), np.float32)
    obs[0] = env.get_obs()
    state = np.zeros((n, 9), np.int32)
    rew = np.zeros(n, np.float64)
    acts = np.zeros(n, np.int64)
    for t in range(n):
        a = int(rng.integers(0, 6))
        if goal_bias > 0 and rng.random() < goal_bias:
            # steer towards the column above the goal so that the goal branch is exercised
            dx, dy, dz = goal[0] - env.x, goal[1] - env.y, goal[2] - env.z
            want = None
            if dx != 0:
                want = (1 if dx > 0 else -1, 0, 0)
            elif dy != 0:
                want = (0, 1 if dy > 0 else -1, 0)
            elif dz != 0:
                a = 4 if dz > 0 else 5
            if want is not None:
                table = {0: [(0, 1, 0), (1, 0, 0), (0, -1, 0), (-1, 0, 0)],
                         1: [(1, 0, 0), (0, -1, 0), (-1, 0, 0), (0, 1, 0)],
                         2: [(0, -1, 0), (-1, 0, 0), (0, 1, 0), (1, 0, 0)],
                         3: [(-1, 0, 0), (0, 1, 0), (1, 0, 0), (0, -1, 0)]}
                for cand in range(4):
                    if table[cand][env.facing] == want:
                        a = cand
        acts[t] = a
        with quiet():
            o, r, term, trunc, _ = env.step(a)
        obs[t + 1] = o
        rew[t] = r
        state[t] = [env.x, env.y, env.z, env.facing, env.visited_count, env.bump_count, env.step_count,
                    int(term), int(trunc)]
    return {
        "room": room, "L": L, "seed": seed, "aseed": aseed, "n": n,
        "start": np.array(start, np.int32), "goal": np.array(goal, np.int32), "actions": acts,
        "state": state, "reward": rew, "obs": obs, "final_ig": env.internal_grid.astype(np.int32),
        "total_free": int(env.total_free_cells),
        "obs_sha": sha16(obs.tobytes()), "state_sha": sha16(state.tobytes()),
        "reward_sha": sha16(rew.astype(np.float32).tobytes()),
    }


def seeded_picks(CubicAgent):
    """reset(seed=s) over a room LIST in an explicit (sorted) order: which room and which start cell."""
    out = {}
    for d in ("P1_training", "P3_training"):
        rooms = sorted((REF / "rooms" / d).glob("*.txt"))
        rows = []
        for s in list(range(0, 24)) + [42, 43, 44, 45, 46, 47, 48, 49, 12345, 2 ** 31 - 1]:
            env = CubicAgent(room_path=str(rooms[0].parent), local_map_length=4)
            env.rooms = list(rooms)
            # find which room was picked by replaying the reference's own first choice()
            pyrandom.seed(s)
            picked = pyrandom.choice(env.rooms)
            with quiet():
                env.reset(seed=s)
            rows.append({"seed": s, "room": picked.name, "room_index": rooms.index(picked),
                         "start": [int(env.x), int(env.y), int(env.z)],
                         "dims": [int(env.width), int(env.depth), int(env.height)]})
        out[d] = {"rooms_sorted": [p.name for p in rooms], "picks": rows}
    return out


def flatten(cases):
    flat = {}
    meta = []
    for i, c in enumerate(cases):
        m = {}
        for k, v in c.items():
            if isinstance(v, np.ndarray):
                flat[f"c{i}_{k}"] = v
            else:
                m[k] = v
        meta.append(m)
    flat["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return flat


def main():
    install_shims()
    from envs.CubicEnv import GridAgent as CubicAgent
    from envs.simpleEnv import GridAgent as SimpleAgent

    rooms = room_answers(CubicAgent, SimpleAgent)
    (HERE / "rooms.json").write_text(json.dumps(rooms, indent=1, sort_keys=True) + "\n")
    print("rooms:", len(rooms))

    E3 = "P1_training/Empty_room_3mx3mx3m_0.25m_cellsize.txt"
    cases = [
        # SURVEY §8c G1..G4
        cubic_trace(CubicAgent, E3, 10, 1, 0, 900),
        cubic_trace(CubicAgent, E3, 4, 1, 0, 900),
        cubic_trace(CubicAgent, "P3_training/kitchen2.txt", 10, 7, 5, 3000),
        cubic_trace(CubicAgent, "P3_training/maze_7x7_seed22.txt", 10, 7, 5, 900),
        # truncation at max_steps = 1000 and stepping on after it
        cubic_trace(CubicAgent, E3, 10, 3, 11, 1100),
        # every P1 room, short
        cubic_trace(CubicAgent, "P1_training/7x7x7_empty_appartment.txt", 10, 5, 1, 600),
        cubic_trace(CubicAgent, "P1_training/7x7x7_empty_room.txt", 10, 6, 2, 600),
        cubic_trace(CubicAgent, "P1_training/Empty_room_2x3.5mx3.5mx3m_with_2x2x3m_connection_0.25m_cellsize.txt", 10, 7, 3, 600),
        cubic_trace(CubicAgent, "P1_training/empty-can-4mx4mx3m.txt", 10, 8, 4, 600),
        cubic_trace(CubicAgent, "P1_evaluate/Empty_room_5mx5mx3m_cellsize_0.25.txt", 10, 9, 5, 600),
        # odd shapes: H = 9, H = 6, open shells, tiny free volumes, other ray lengths
        cubic_trace(CubicAgent, "P2_training/tightcorridor.txt", 10, 2, 6, 700),
        cubic_trace(CubicAgent, "P3_training/maze_3d_tunnels.txt", 3, 2, 7, 500),
        cubic_trace(CubicAgent, "P3_training/maze_dead1_11.txt", 1, 4, 8, 500),
        cubic_trace(CubicAgent, "P2_training/maze_8x8_seed22.txt", 15, 4, 9, 500),
        cubic_trace(CubicAgent, "P2_training/small_bedroom.txt", 7, 4, 10, 500, crash_penalty=-0.5),
        # termination (>= 84 % explored) under a coverage policy, then 20 further steps
        cubic_trace(CubicAgent, E3, 10, 1, 21, 960, policy="greedy"),
        cubic_trace(CubicAgent, "P2_training/tightcorridor.txt", 10, 1, 22, 320, policy="greedy"),
        cubic_trace(CubicAgent, "P3_training/maze_3d_tunnels.txt", 10, 1, 23, 170, policy="greedy"),
    ]
    for c in cases:
        term = int(c["state"][:, 7].max())
        trunc = int(c["state"][:, 8].max())
        print(f"cubic {c['room']:<60s} L={c['L']:<2d} n={c['n']:<4d} obs={c['obs_sha']} state={c['state_sha']} "
              f"rew={c['reward_sha']} sumR={c['reward'].sum():.2f} term={term} trunc={trunc} maxIG={c['final_ig'].max()}")
    np.savez_compressed(HERE / "cubic_traces.npz", **flatten(cases))

    scases = [
        simple_trace(SimpleAgent, E3, 4, 1, 0, 600),
        simple_trace(SimpleAgent, "P3_training/kitchen2.txt", 4, 7, 5, 600),
        simple_trace(SimpleAgent, "P3_training/maze_7x7_seed22.txt", 4, 7, 5, 600),
        simple_trace(SimpleAgent, "P2_training/tightcorridor.txt", 10, 2, 6, 400),
        simple_trace(SimpleAgent, E3, 4, 3, 11, 1100),
        simple_trace(SimpleAgent, E3, 4, 5, 12, 400, goal_bias=0.7),
        simple_trace(SimpleAgent, "P1_training/empty-can-4mx4mx3m.txt", 6, 5, 13, 600, goal_bias=0.7),
    ]
    for c in scases:
        print(f"simple {c['room']:<60s} L={c['L']:<2d} n={c['n']:<4d} obs={c['obs_sha']} state={c['state_sha']} "
              f"rew={c['reward_sha']} sumR={c['reward'].sum():.2f} term={int(c['state'][:,7].max())}")
    np.savez_compressed(HERE / "simple_traces.npz", **flatten(scases))

    (HERE / "seeded_picks.json").write_text(json.dumps(seeded_picks(CubicAgent), indent=1) + "\n")


if __name__ == "__main__":
    main()
