"""Edge-case rooms shared by the CPU (emulated) and GPU parity tests: the engine's size limits, degenerate rooms, rooms
whose boundary is open at the highest bit positions of the packed occupancy words."""
import numpy as np

from nav3d.room_tools import generate_room
from nav3d.rooms import rooms_from_grids


def max_size_rooms():
    """64x64x16 (the limit of include/nav3d.h): furnished and closed; and a copy with holes in the +x, +y, +z faces so that
    rays and moves reach x = 63 (bit 63 of the u64 row words), y = 63 and z = 15 (bit 15 of the column words)."""
    closed = generate_room("furnished", (64, 64, 16), seed=1)
    holes = closed.copy()
    holes[63, 20:40, 3:12] = 0
    holes[10:50, 63, 2:14] = 0
    holes[5:60, 5:60, 15] = 0
    holes[0, 30:34, 5:9] = 0
    holes[12:20, 0, 4:8] = 0
    holes[30:40, 30:40, 0] = 0
    return rooms_from_grids([closed, holes])


def degenerate_rooms():
    """3x3x3 with a single free cell (every move bumps; the first step both terminates and truncates), a 3x64x3 corridor,
    a 64x3x16 slab, and a two-cell room for counter saturation."""
    def shell(w, d, h):
        g = np.zeros((w, d, h), dtype=np.int8)
        g[0], g[-1], g[:, 0], g[:, -1], g[:, :, 0], g[:, :, -1] = -2, -2, -2, -2, -2, -2
        return g
    return rooms_from_grids([shell(3, 3, 3), shell(3, 64, 3), shell(64, 3, 16), shell(3, 3, 4)])


def thin_open_rooms():
    """All-free rooms without any wall (every boundary cell is reachable: out-of-bounds on all six sides), incl. 1-cell-thick."""
    return rooms_from_grids([np.zeros((5, 4, 3), dtype=np.int8), np.zeros((3, 3, 16), dtype=np.int8),
                             np.zeros((64, 3, 3), dtype=np.int8)])
