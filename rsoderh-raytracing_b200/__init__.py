"""B200-native backend for the per-pixel path-tracing hot path of rsoderh-raytracing.

The directory name carries a hyphen (it mirrors the reference repository's name), so it
is imported under the module name `rsoderh_raytracing_b200` through `_pkg.load()` at
the repo root.
"""
from .api import (REFERENCE_MAX_BOUNCES, Camera, Environment, EnvironmentMaps, RtError, Scene,
                  DeviceBvhBuilder, SceneBuilder, State, StateGroup, pinned_array, write_pfm, write_png)
from . import _ffi, samplesplit

__all__ = ["REFERENCE_MAX_BOUNCES", "Camera", "Environment", "EnvironmentMaps", "RtError",
           "Scene", "DeviceBvhBuilder", "SceneBuilder", "State", "StateGroup", "pinned_array", "write_pfm", "write_png", "_ffi", "samplesplit"]
