"""Puts the product package on sys.path: ``3d-navigation-reinforcement-learning_b200/`` is not an importable name, the
Python package inside it (``nav3d``) is."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "3d-navigation-reinforcement-learning_b200"
if str(PKG_DIR) not in sys.path:
    sys.path.insert(0, str(PKG_DIR))
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
