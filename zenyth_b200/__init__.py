"""zenyth_b200 — B200-native scene-update path for the Zenyth engine (TNtube/Zenyth).

The product is zenyth_b200/lib/libzenyth_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/zenyth_b200.h).  This package is the Python host mirror of the engine-idiom API; it loads
that library with ctypes and fails loudly when it is missing.  There is no CPU fallback.
"""
from .engine import Camera, CameraDesc, GpuContext, Mesh, Scene, SortHierarchy, Transform, ZenythError, ZN_NO_PARENT, mat4

__all__ = ["Camera", "CameraDesc", "GpuContext", "Mesh", "Scene", "SortHierarchy", "Transform", "ZenythError", "ZN_NO_PARENT", "mat4"]
__version__ = "0.1.0"
