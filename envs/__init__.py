"""Drop-in ``envs`` package: the module names the reference's drivers import (``envs.CubicEnv``, ``envs.simpleEnv`` and
``envs.Venv`` — the last is imported by the reference's ``train/evaluate_grid.py:4`` but missing from its tree), backed by
the CUDA engine in ``3d-navigation-reinforcement-learning_b200/``."""
import sys
from pathlib import Path

_root = Path(__file__).resolve().parent.parent
if str(_root) not in sys.path:
    sys.path.insert(0, str(_root))
import _nav3d_path  # noqa: E402,F401
