"""Pins oracle/py_cubic.py (the faithful-cost Python port timed by bench.py's reference arm) to the reference traces."""
import numpy as np

from conftest import ROOMS
from nav3d.rooms import load_room_file
from oracle.py_cubic import PyCubic, free_cells


def test_py_port_replays_reference_traces(cubic_traces):
    for c in cubic_traces[:5] + cubic_traces[10:12] + cubic_traces[-2:]:
        room = load_room_file(ROOMS / c["room"])
        g = room.grid.astype(int)
        env = PyCubic(c["L"], c["crash_penalty"])
        fc = free_cells(g)
        assert len(fc) == c["total_free"] and np.array_equal(fc, room.free_cells())
        obs = env.reset(g, len(fc), c["start"])
        assert np.array_equal(obs.view(np.uint32), c["obs"][0].view(np.uint32))
        n = min(c["n"], 400) if c["policy"] == "random" else c["n"]
        for t in range(n):
            o, r, term, trunc = env.step(int(c["actions"][t]))
            got = [env.x, env.y, env.z, env.facing, env.visited, env.bumps, env.steps, int(term), int(trunc)]
            assert got == list(c["state"][t]), (c["room"], t)
            assert r == c["reward"][t], (c["room"], t)
            assert np.array_equal(o.view(np.uint32), c["obs"][t + 1].view(np.uint32)), (c["room"], t)
