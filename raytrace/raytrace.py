"""``raytrace.raytrace`` -> the B200 implementation (ray_trace_pb_b200.raytrace)."""
from ray_trace_pb_b200.raytrace import *  # noqa: F401,F403
from ray_trace_pb_b200.raytrace import __all__  # noqa: F401
