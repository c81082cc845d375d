"""``envs.Venv`` — the module name ``train/evaluate_grid.py:4`` of the reference imports (absent from the reference tree);
it is the CubicEnv ``GridAgent``."""
from .CubicEnv import FINISH_PERCENTAGE, NN_SIZES, GridAgent  # noqa: F401
