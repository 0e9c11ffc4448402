"""flexpart_b200 -- B200-native engine for FLEXPART's per-particle timestep.

The product is the CUDA library ``libfpb.so`` (C ABI: ``include/fpb.h``) plus
the host companion ``libfpb_host.so`` (``include/fpb_host.h``).  This package
is the thin ctypes binding a Python host uses; the Fortran binding is given in
INTEGRATION.md.  There is no CPU fallback: engine calls fail loudly when the
CUDA library or a device is missing.
"""
from .abi import (FpbConfig, FpbMetPtrs, FpbParticlePtrs, FpbStepStats, load_engine_lib,
                  load_host_lib, FpbError, RNG_REFERENCE, RNG_PHILOX_INDEX, RNG_PHILOX,
                  MATH_FAST, MATH_STRICT, SCATTER_ATOMIC, SCATTER_DETERMINISTIC, ITRA_DEAD)
from .host import (make_config, MetFields, Particles, synth_heights, Releases, RunSpec,
                   timemanager, release_particles, ReleaseState, outgrid_geometry,
                   verttransform_heights, synth_hybrid_levels, synth_rawmet)
from .engine import Engine

__all__ = [
    "FpbConfig", "FpbMetPtrs", "FpbParticlePtrs", "FpbStepStats", "FpbError", "Engine",
    "make_config", "MetFields", "Particles", "synth_heights", "Releases", "RunSpec",
    "timemanager", "release_particles", "ReleaseState", "outgrid_geometry", "verttransform_heights", "synth_hybrid_levels", "synth_rawmet", "load_engine_lib", "load_host_lib",
    "RNG_REFERENCE", "RNG_PHILOX_INDEX", "RNG_PHILOX", "MATH_FAST", "MATH_STRICT",
    "SCATTER_ATOMIC", "SCATTER_DETERMINISTIC", "ITRA_DEAD",
]
