"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/nav3d.h declares,
and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import re

import pytest

from conftest import ROOT, has_cuda


def header_symbols():
    text = (ROOT / "include" / "nav3d.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nav3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build_lib()
    from nav3d import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nav3d.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "nav3d/_lib.py binds exactly the header's entry points"
    assert lib.nav3d_abi_version() == 2


def test_config_struct_layout_matches_header():
    from nav3d._lib import Config, RoomDesc
    assert ctypes.sizeof(Config) == 56 and Config.seed.offset == 32 and Config.crash_penalty.offset == 40
    assert ctypes.sizeof(RoomDesc) == 40 and RoomDesc.grid.offset == 16 and RoomDesc.has_start.offset == 24
    from nav3d._lib import RewardParams
    assert ctypes.sizeof(RewardParams) == 80 and RewardParams.crash_penalty.offset == 24


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu():
    import nav3d
    from nav3d._lib import Config
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nav3d.Engine(4, [nav3d.default_box_room()])
    lib = nav3d._lib.load()
    cfg = Config(abi_version=2, device=0, n_envs=4, env_kind=0, local_map_length=4, auto_reset=1, lanes_per_env=0,
                 env_id0=0, seed=0, crash_penalty=-2.0, cell_size=0.25)
    h = ctypes.c_void_p()
    rc = lib.nav3d_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -3 and b"cuda" in lib.nav3d_last_error().lower()     # NAV3D_ERR_CUDA, nothing was created
    assert not h


def test_argument_validation_needs_no_gpu():
    import nav3d
    lib = nav3d._lib.load()
    from nav3d._lib import Config
    h = ctypes.c_void_p()
    bad = Config(abi_version=99, device=0, n_envs=4, env_kind=0, local_map_length=4)
    assert lib.nav3d_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    bad = Config(abi_version=2, device=0, n_envs=4, env_kind=0, local_map_length=0)
    assert lib.nav3d_create(ctypes.byref(bad), ctypes.byref(h)) == -2
    assert lib.nav3d_create(None, ctypes.byref(h)) == -1
    # the reward constants of compute_reward (CubicEnv.py:175-221) are the struct's defaults
    from nav3d._lib import RewardParams
    p = RewardParams()
    lib.nav3d_reward_params_default(ctypes.byref(p))
    assert [getattr(p, f) for f, _ in RewardParams._fields_] == [-0.05, 0.02, 0.5, -2.0, 0.15, 0.05, 0.5, 1.0, 100.0, -5.0]
