"""mcbrat3d_b200 -- B200-native photon-tracing hot path of MCBRaT3D.

Host-side modules mirror the reference's Fortran modules for this path (same names,
argument meaning and error behaviour); the photon loop itself runs in hand-written
sm_100a CUDA kernels behind the C ABI declared in ``include/mcbrat_cuda.h``
(``csrc/libmcbrat_cuda.so``).  There is no CPU fallback: the integrator raises if the
CUDA library is missing.
"""
__version__ = "0.1.0"
