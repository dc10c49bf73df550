"""Stub of pytransit (oracle only): exposes the restated QuadraticModel (oracle/quadmodel.py)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.abspath(os.path.join(_here, "..", "..", ".."))
if _repo not in sys.path:
    sys.path.insert(0, _repo)
from oracle.quadmodel import QuadraticModel  # noqa: E402,F401
