"""Shim: the oracle's ctypes binding lives next to the oracle (oracle/binding.py)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from binding import *  # noqa: F401,F403
from binding import OracleFit, build, dp, lib  # noqa: F401
