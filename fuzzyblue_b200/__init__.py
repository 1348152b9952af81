"""fuzzyblue_b200 — B200-native (sm_100a) atmosphere LUT precompute and sky evaluation.

Same six public names as the reference crate (/root/reference/src/lib.rs:8-12); see ``api.py``.
"""
from .api import (Atmosphere, Builder, DensityProfile, DensityProfileLayer, DrawParameters, FuzzyblueError, Parameters,
                  PendingAtmosphere, Renderer, build_batch, precompute_host, sky_radiance, sun_and_sky_irradiance)

__all__ = ["Atmosphere", "Builder", "Parameters", "PendingAtmosphere", "DrawParameters", "Renderer", "DensityProfile",
           "DensityProfileLayer", "FuzzyblueError", "build_batch", "precompute_host", "sky_radiance", "sun_and_sky_irradiance"]
