"""
ray_trace_pb_b200 -- B200-native batch ray tracer behind the ``raytrace.raytrace`` / ``raytrace.materials`` API of
QI2lab/ray_trace_pb.

    import ray_trace_pb_b200.raytrace as rt
    import ray_trace_pb_b200.materials as rtm

(or, unchanged user code: ``import raytrace.raytrace as rt`` through the ``raytrace`` shim package at the repo root).

``raytrace``   host objects + the drop-in calls          ``materials``  glass catalogue
``device``     device-resident API (torch tensors as HBM buffers, ray sources, fused reductions)
``sharding``   ray-range sharding over the GPUs of one node and the NCCL all-reduce of reduced products
``engine``     prescription packing                       ``_ffi``       ctypes binding of librtb.so (include/rtb.h)
``analysis``   spot / pupil / focus sweeps on the device  ``persist``    sweep results in the scripts' zarr layout
"""
__version__ = "0.1.0"
