"""Pins the CPU oracle (oracle/nav3d_oracle.c) and the room parser to the UNMODIFIED reference.

The fixtures under tests/golden were produced by tests/golden/make_golden.py, which imports the reference's own
envs/CubicEnv.py and envs/simpleEnv.py (only possible in the build container)."""
import hashlib

import numpy as np
import pytest

from conftest import ROOMS
from nav3d.rooms import load_room_file


def sha16(b):
    return hashlib.sha256(b).hexdigest()[:16]


def test_room_parser_matches_reference_grids(rooms_json):
    assert len(rooms_json) == 60
    for key, ans in rooms_json.items():
        r = load_room_file(ROOMS / key)
        assert list(r.dims) == ans["dims"], key
        assert sha16(r.grid.tobytes()) == ans["cubic_grid_sha"], key
        assert int((r.grid == -2).sum()) == ans["n_wall_cells"], key
        assert len(r.free_cells()) == ans["total_free_cells"], key
        rs = load_room_file(ROOMS / key, simple=True)
        assert sha16(rs.grid.tobytes()) == ans["simple_grid_sha"], key
        assert len(rs.free_cells()) == ans["simple_total_free_cells"], key


def test_oracle_free_cells_match_parser(rooms_json, oracle):
    for key in list(rooms_json)[::7]:
        r = load_room_file(ROOMS / key)
        o = oracle.OracleRoom(r.grid, -2)
        fc = r.free_cells()
        assert o.n_free == len(fc) == rooms_json[key]["total_free_cells"]
        for k in (0, len(fc) // 2, len(fc) - 1):
            assert o.free_cell(k) == tuple(fc[k])


def test_cubic_oracle_replays_reference_traces(cubic_traces, oracle):
    assert len(cubic_traces) >= 18
    for c in cubic_traces:
        room = load_room_file(ROOMS / c["room"])
        oroom = oracle.OracleRoom(room.grid, -2)
        assert oroom.n_free == c["total_free"]
        env = oracle.OracleCubic(c["L"], c["crash_penalty"])
        n = c["n"]
        obs = np.zeros((n + 1, 80), np.float32)
        state = np.zeros((n, 9), np.int32)
        rew = np.zeros(n, np.float64)
        flags = np.zeros((n, 3), np.int8)
        obs[0] = env.reset(oroom, c["start"])
        for t in range(n):
            o, r, term, trunc = env.step(int(c["actions"][t]))
            s = env.state()
            obs[t + 1] = o
            rew[t] = r
            state[t] = [s[0], s[1], s[2], s[3], s[4], s[5], s[6], int(term), int(trunc)]
            flags[t] = [s[7], s[8], s[9]]
        tag = f"{c['room']} L={c['L']} policy={c['policy']}"
        assert np.array_equal(state, c["state"]), tag
        assert np.array_equal(flags, c["flags"]), tag
        assert np.array_equal(rew, c["reward"]), tag               # f64, bit-exact
        assert np.array_equal(obs.view(np.uint32), c["obs"].view(np.uint32)), tag   # f32, bit-exact
        assert np.array_equal(env.grid().astype(np.int32), c["final_ig"]), tag
        assert sha16(obs.tobytes()) == c["obs_sha"] and sha16(state.tobytes()) == c["state_sha"]
        assert sha16(rew.astype(np.float32).tobytes()) == c["reward_sha"]


def test_survey_golden_hashes(cubic_traces):
    """SURVEY.md §8c G1..G4 (hashes captured by the surveyor from the unmodified reference)."""
    want = {("P1_training/Empty_room_3mx3mx3m_0.25m_cellsize.txt", 10, 900): ("f5c361db16d3e3ed", "2b8b37e46228b86c", "5d39bf549ce0065d"),
            ("P1_training/Empty_room_3mx3mx3m_0.25m_cellsize.txt", 4, 900): ("ed4960ac580a300d", "2b8b37e46228b86c", "5d39bf549ce0065d"),
            ("P3_training/kitchen2.txt", 10, 3000): ("735a835271e15711", "602723c252edcaed", "f19a2d4dafbbfcb5"),
            ("P3_training/maze_7x7_seed22.txt", 10, 900): ("ab1161a38694f938", "522561f54a048a4e", "9416dec5069c1b5b")}
    seen = 0
    for c in cubic_traces:
        k = (c["room"], c["L"], c["n"])
        if k in want and c["policy"] == "random" and c["aseed"] in (0, 5):
            assert (c["obs_sha"], c["state_sha"], c["reward_sha"]) == want[k]
            seen += 1
    assert seen == 4


def test_simple_oracle_replays_reference_traces(simple_traces, oracle):
    assert len(simple_traces) >= 7
    for c in simple_traces:
        room = load_room_file(ROOMS / c["room"], simple=True)
        oroom = oracle.OracleRoom(room.grid, 2)
        assert oroom.n_free == c["total_free"]
        env = oracle.OracleSimple(c["L"], 0.25)
        n = c["n"]
        dim = 6 * c["L"] + 7
        obs = np.zeros((n + 1, dim), np.float32)
        state = np.zeros((n, 9), np.int32)
        rew = np.zeros(n, np.float64)
        env.reset(oroom, c["start"], c["goal"])
        obs[0] = env.get_obs()
        for t in range(n):
            o, r, term, trunc = env.step(int(c["actions"][t]))
            s = env.state()
            obs[t + 1] = o
            rew[t] = r
            state[t] = [s[0], s[1], s[2], s[3], s[4], s[5], s[6], int(term), int(trunc)]
        tag = f"{c['room']} L={c['L']}"
        assert np.array_equal(state, c["state"]), tag
        assert np.array_equal(rew, c["reward"]), tag
        assert np.array_equal(obs.view(np.uint32), c["obs"].view(np.uint32)), tag
        assert np.array_equal(env.grid().astype(np.int32), c["final_ig"]), tag


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors in the Random123 distribution)."""
    assert list(oracle.philox4x32_10([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
