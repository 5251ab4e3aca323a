"""gi_raytracer_b200 — a B200-native (sm_100a) rendering hot path behind GI_Raytracer's scene API.

Layout: csrc/ holds the CUDA kernels and the C ABI (include/gi_api.h -> libgi_b200.so) plus the host-side C++
mirror of the reference's scene classes (libgi_host.so); this Python package is the thin harness over both
(ctypes), used by the tests, bench.py and the multi-GPU driver.  There is no CPU fallback: without the CUDA
library the compute calls raise.
"""
__version__ = "0.1.0"
