#!/usr/bin/env python
"""Golden text renderings of the UNMODIFIED reference env (envs/CubicEnv.py:475-499, render_mode="human"), written to
tests/golden/render_text.json.  Run in the container that has /root/reference:  python tests/golden/make_render_golden.py
The case: P1_training/Empty_room_3mx3mx3m (12x12x12), L = 10, reset(seed=1), actions from default_rng(0); the printed
frame after the reset and after steps 10, 40 and 120 (own cell, known walls, counted, seen and unknown cells all appear)."""
import contextlib
import io
import json
from pathlib import Path

import numpy as np

from make_golden import REF, install_shims

HERE = Path(__file__).resolve().parent
ROOM = "P1_training/Empty_room_3mx3mx3m_0.25m_cellsize.txt"


def main():
    install_shims()
    from envs.CubicEnv import GridAgent
    env = GridAgent(room_path=str(REF / "rooms" / "P1_training"), local_map_length=10, render_mode="human")
    env.rooms = [REF / "rooms" / ROOM]
    frames = {}

    def grab(tag):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            env.render()
        frames[tag] = buf.getvalue()

    with contextlib.redirect_stdout(io.StringIO()):
        env.reset(seed=1)
    grab("0")
    actions = np.random.default_rng(0).integers(0, 6, size=120)
    for t, a in enumerate(actions, 1):
        with contextlib.redirect_stdout(io.StringIO()):
            env.step(int(a))
        if t in (10, 40, 120):
            grab(str(t))
    (HERE / "render_text.json").write_text(json.dumps({"room": ROOM, "L": 10, "seed": 1, "action_seed": 0, "frames": frames},
                                                      indent=1) + "\n")
    print({k: len(v) for k, v in frames.items()})


if __name__ == "__main__":
    main()
