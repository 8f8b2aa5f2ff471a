"""Parallel runtime: topology, process groups, TP/SP layers, PP schedules, DP buffers."""
from . import state as parallel_state  # noqa: F401
from .topology import Topology, layer_offset  # noqa: F401
