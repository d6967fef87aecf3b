"""``raytrace.materials`` -> the B200 implementation (ray_trace_pb_b200.materials)."""
from ray_trace_pb_b200.materials import *  # noqa: F401,F403
from ray_trace_pb_b200.materials import __all__  # noqa: F401
