"""B200-native drop-in for the reference package `threedgrt_tracer` (threedgrt_tracer/__init__.py)."""
from .tracer import Tracer, OptixTracer  # noqa: F401
