"""vbs_b200: B200-native marker pipeline (tracking -> 3D displacement -> plane tilt).

Drop-in for the per-frame hot path of UPM-ROB-Lab/Vision-basedSensor.  Host code is
Python; all arithmetic on the path runs in hand-written sm_100a CUDA kernels behind
the C ABI declared in ``include/vbs.h`` (``csrc/libvbs_b200.so``).  There is no CPU
fallback: importing :mod:`vbs_b200.capi` raises if the library is missing.
"""
__all__ = ["synth"]
__version__ = "0.1.0"
