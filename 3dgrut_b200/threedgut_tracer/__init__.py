"""B200-native drop-in for the reference package `threedgut_tracer` (threedgut_tracer/__init__.py)."""
from .tracer import Tracer, SplatRaster, ShutterType, SensorPose3D, fromOpenCVPinholeCameraModelParameters  # noqa: F401
