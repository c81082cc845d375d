import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _nav3d_path  # noqa: E402,F401  (puts the nav3d package on sys.path)

GOLDEN = Path(__file__).resolve().parent / "golden"
ROOMS = ROOT / "rooms"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_traces(name):
    d = np.load(GOLDEN / name)
    meta = json.loads(bytes(d["meta"]).decode())
    cases = []
    for i, m in enumerate(meta):
        c = dict(m)
        for k in d.files:
            if k.startswith(f"c{i}_"):
                c[k[len(f"c{i}_"):]] = d[k]
        cases.append(c)
    return cases


@pytest.fixture(scope="session")
def cubic_traces():
    return load_traces("cubic_traces.npz")


@pytest.fixture(scope="session")
def simple_traces():
    return load_traces("simple_traces.npz")


@pytest.fixture(scope="session")
def rooms_json():
    return json.loads((GOLDEN / "rooms.json").read_text())


@pytest.fixture(scope="session")
def oracle():
    from oracle import c_oracle
    c_oracle.lib()
    return c_oracle


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False
