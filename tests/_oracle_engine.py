"""Test helper: the CPU oracle port of the engine (oracle/engine_port.py) under its test name."""
from oracle.engine_port import OracleEngine, install, uninstall  # noqa: F401
