"""Host-side logic: room parser quirks, the integer form of the 84 % finish test, seeded picks, spaces."""
import json
import random

import numpy as np
import pytest

from conftest import GOLDEN, ROOMS
from nav3d.rooms import default_box_room, list_room_files, load_room_dir, load_room_file, parse_room_text
from nav3d.spaces import cubic_spaces, simple_spaces


def test_finish_threshold_integer_form():
    """visited/total >= 0.84 in f64 (CubicEnv.py:212-213) == 25*visited >= 21*total for every total the engine allows."""
    t = np.arange(1, 65537, dtype=np.int64)
    v_int = -(-21 * t // 25)                      # smallest v with 25 v >= 21 t
    assert np.all(v_int.astype(np.float64) / t >= 0.84)
    assert np.all((v_int - 1).astype(np.float64) / t < 0.84)


def test_parser_quirks():
    txt = "Size=3,4,3\nLayer z=0\n2 2 2\n2 2 2\n2 2 2\n2 2 2\nLayer z=1\n2 2 2\n2 0 2\n2 -2 2\n2 2 2\nLayer z=-1\n2 2 2\n2 2 2\n2 2 2\n2 2 2\n"
    r = parse_room_text(txt)
    assert r.dims == (3, 4, 3) and r.grid[1, 1, 1] == 0 and r.grid[1, 2, 1] == -2 and (r.grid[:, :, 2] == -2).all()
    rs = parse_room_text(txt, simple=True)
    assert rs.grid[1, 2, 1] == -2 and rs.wall_code == 2 and len(rs.free_cells()) == 2    # -2 is free for simpleEnv
    with pytest.raises(ValueError, match="has 2 values, but width is 3"):
        parse_room_text("Size=3,1,1\nLayer z=0\n2 2\n")
    with pytest.raises(IndexError):
        parse_room_text("Size=1,1,1\nLayer z=0\n2\n2\n")
    with pytest.raises(IndexError):
        parse_room_text("Size=1,1,1\nLayer z=5\n2\n")
    r = parse_room_text("Start position=1,2,3\nGoal=3,2,1\nSize=1,1,1\nLayer z=0\n0\n")
    assert r.start == (1, 2, 3) and r.goal == (3, 2, 1)


def test_kitchen2_negative_layer_and_wall_codes():
    r = load_room_file(ROOMS / "P3_training" / "kitchen2.txt")
    assert (r.grid[:, :, 2] == 0).all()          # the real layer 2 is never written (the file says z=-2)
    assert len(r.free_cells()) == 6414           # tests/golden/rooms.json (from the unmodified reference)
    rs = load_room_file(ROOMS / "P3_training" / "kitchen2.txt", simple=True)
    assert len(rs.free_cells()) == 6760          # this file writes walls as -2, which simpleEnv does not treat as walls


def test_room_dir_order_is_glob_order():
    files = list_room_files(ROOMS / "P1_training")
    assert len(files) == 5 and sorted(files) == sorted((ROOMS / "P1_training").glob("*.txt"))
    assert [r.name for r in load_room_dir(ROOMS / "P1_training")] == [p.name for p in files]


def test_default_box():
    r = default_box_room()
    assert r.dims == (20, 20, 12) and len(r.free_cells()) == 18 * 18 * 10


def test_seeded_picks_follow_cpython_random():
    """reset(seed=s): random.seed(s); room = random.choice(rooms); start = random.choice(possible_start_pose)."""
    picks = json.loads((GOLDEN / "seeded_picks.json").read_text())
    for d, blob in picks.items():
        rooms = [load_room_file(ROOMS / d / n) for n in blob["rooms_sorted"]]
        cells = [r.free_cells() for r in rooms]
        for p in blob["picks"]:
            rng = random.Random(p["seed"])
            ri = rng.choice(range(len(rooms)))
            k = rng.choice(range(len(cells[ri])))
            assert ri == p["room_index"] and list(cells[ri][k]) == p["start"], (d, p["seed"])


def test_spaces():
    a, o = cubic_spaces()
    assert a.n == 6 and o.shape == (80,) and o.dtype == np.float32
    a, o = simple_spaces(4)
    assert o.shape == (31,)


def test_marking_list_capacity_bound():
    """csrc/nav3d_core.cuh mark_tasks_per_lane: a warp's shared-memory marking list holds 32 x [(min(2L+1, W) + 6) / 4 +
    (min(2L+1, D) + 6) / 4 + 4] tasks.  Brute force over every agent position: an x (y) run of the cells within L of the
    agent never spans more tiles of the bordered volume (2 border columns, 4 columns per tile) than the bound allows."""
    def bound(L, w):
        return (min(2 * L + 1, w) + 6) // 4

    for L in list(range(1, 34)) + [40, 100, 255]:
        for w in range(3, 65):
            worst = 0
            for x in range(w):
                f0, f1 = max(0, x - L), min(w - 1, x + L)
                worst = max(worst, ((f1 + 2) >> 2) - ((f0 + 2) >> 2) + 1)
            assert worst <= bound(L, w), (L, w, worst)
