"""nav3d — B200-native batched voxel-navigation environments (drop-in for the reference's envs/CubicEnv.py hot path).

Importing this package loads ``lib/libnav3d_b200.so``; there is no CPU fallback."""
from . import _lib
from ._lib import Nav3dError
from .rooms import (Room, default_box_room, list_room_files, load_room_dir, load_room_file, parse_room_text,
                    rooms_from_grids)
from .spaces import Box, Discrete, cubic_spaces, simple_spaces

_lib.load()

from .engine import EPISODE_DTYPE, STATE_FIELDS, Engine  # noqa: E402
from .vec_env import BatchedCubicEnv, BatchedSimpleEnv, NumpyVecEnv, StepInfo  # noqa: E402

__all__ = ["Engine", "BatchedCubicEnv", "BatchedSimpleEnv", "NumpyVecEnv", "StepInfo", "Room", "Nav3dError", "parse_room_text", "load_room_file",
           "load_room_dir", "list_room_files", "default_box_room", "rooms_from_grids", "Discrete", "Box",
           "cubic_spaces", "simple_spaces", "EPISODE_DTYPE", "STATE_FIELDS"]
