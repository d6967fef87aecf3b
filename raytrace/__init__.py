"""Drop-in name for user code written against QI2lab/ray_trace_pb: ``import raytrace.raytrace as rt``."""
