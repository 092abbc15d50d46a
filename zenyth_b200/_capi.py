"""ctypes binding of include/zenyth_b200.h — one prototype per exported symbol.

The library is the product; there is no Python or CPU fallback.  If libzenyth_b200.so is missing
`load()` raises with the build command instead of degrading.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libzenyth_b200.so")

ZN_OK = 0
ZN_ERR_INVALID = 1
ZN_ERR_CUDA = 2
ZN_ERR_NO_DEVICE = 3
ZN_ERR_STATE = 4
ZN_ERR_NOMEM = 5
ZN_NO_PARENT = 0xFFFFFFFF

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)
vp = C.c_void_p


class SceneDesc(C.Structure):
    _fields_ = [("entity_count", C.c_uint32), ("level_count", C.c_uint32), ("level_offsets", vp)]


class SceneBuffers(C.Structure):
    _fields_ = [("world", vp), ("mvp", vp), ("visible", vp), ("visible_count", vp), ("visibility_record", vp),
                ("visibility_record_words", C.c_uint32), ("entity_count", C.c_uint32)]


class FrameHostIO(C.Structure):
    _fields_ = [("position", vp), ("rotation", vp), ("scale", vp), ("world", vp), ("mvp", vp), ("visible", vp),
                ("visible_count", C.c_uint32), ("first", C.c_uint32), ("count", C.c_uint32)]


class CullOutput(C.Structure):
    _fields_ = [("mvp_dev", vp), ("visible_dev", vp), ("visible_capacity", C.c_uint32), ("visible_count_dev", vp)]


class MeshDesc(C.Structure):
    _fields_ = [("vertex_count", C.c_uint32), ("instance_count", C.c_uint32), ("instance_first_vertex", vp)]


class MeshBuffers(C.Structure):
    _fields_ = [("position_in", vp), ("normal_in", vp), ("position_out", vp), ("normal_out", vp), ("vertex_count", C.c_uint32)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, vp, vp, vp, C.c_size_t)


class ExchangeDesc(C.Structure):
    _fields_ = [("world_size", C.c_uint32), ("rank", C.c_uint32), ("records_per_rank", C.c_uint32), ("record_words", C.c_uint32),
                ("list_capacity", C.c_uint32), ("slots", C.c_uint32), ("timeout_ms", C.c_uint32), ("allgather", ALLGATHER_FN), ("user", vp), ("flags", C.c_uint32)]


# symbol -> (restype, argtypes).  tests/test_abi.py checks this table against include/zenyth_b200.h.
PROTOTYPES = {
    "zn_abi_version": (C.c_int, []),
    "zn_last_error": (C.c_char_p, []),
    "zn_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "zn_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "zn_ctx_destroy": (C.c_int, [vp]),
    "zn_ctx_set_stream": (C.c_int, [vp, vp]),
    "zn_ctx_get_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "zn_ctx_sync": (C.c_int, [vp]),
    "zn_ctx_sm_count": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "zn_host_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "zn_host_free": (C.c_int, [vp, vp]),
    "zn_device_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "zn_device_free": (C.c_int, [vp, vp]),
    "zn_device_read": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "zn_device_write": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "zn_ipc_export": (C.c_int, [vp, vp, vp]),
    "zn_ipc_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "zn_ipc_close": (C.c_int, [vp, vp]),
    "zn_mat4_look_at": (None, [vp, vp, vp, vp]),
    "zn_mat4_perspective": (None, [C.c_float, C.c_float, C.c_float, C.c_float, vp]),
    "zn_mat4_orthographic": (None, [C.c_float] * 6 + [vp]),
    "zn_mat4_perspective_reverse_z": (None, [C.c_float, C.c_float, C.c_float, vp]),
    "zn_mat4_mul": (None, [vp, vp, vp]),
    "zn_frustum_planes": (None, [vp, vp]),
    "zn_hierarchy_sort": (C.c_int, [vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.POINTER(C.c_uint32)]),
    "zn_scene_create": (C.c_int, [vp, C.POINTER(SceneDesc), C.POINTER(vp)]),
    "zn_scene_destroy": (C.c_int, [vp]),
    "zn_scene_set_transforms": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_size_t]),
    "zn_scene_set_bounds": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_size_t]),
    "zn_scene_set_parents": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
    "zn_scene_set_camera": (C.c_int, [vp, vp, vp]),
    "zn_scene_bind_visible_output": (C.c_int, [vp, vp, C.c_uint32, vp]),
    "zn_scene_bind_visibility_record": (C.c_int, [vp, vp]),
    "zn_scene_bind_peer_records": (C.c_int, [vp, vp, C.c_uint32]),
    "zn_scene_expand_visible": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, C.c_uint32, vp]),
    "zn_scene_bind_peer_signal": (C.c_int, [vp, vp, C.c_uint32]),
    "zn_scene_wait_flags": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "zn_scene_set_signal_value": (C.c_int, [vp, C.c_uint32]),
    "zn_scene_expand_visible_signalled": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, vp, vp, C.c_uint32, vp,
                                                    vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "zn_scene_set_expand_mode": (C.c_int, [vp, C.c_int]),
    "zn_scene_set_level_fusion": (C.c_int, [vp, C.c_int]),
    "zn_scene_set_index_base": (C.c_int, [vp, C.c_uint32]),
    "zn_scene_update": (C.c_int, [vp]),
    "zn_scene_update_world": (C.c_int, [vp]),
    "zn_scene_cull": (C.c_int, [vp, vp, vp, C.POINTER(CullOutput)]),
    "zn_scene_clear_camera": (C.c_int, [vp]),
    "zn_scene_set_ticket_origin": (C.c_int, [vp, C.c_uint32]),
    "zn_scene_cull_times": (C.c_int, [vp, c_float_p, c_float_p]),
    "zn_scene_visible_count": (C.c_int, [vp, c_u32_p]),
    "zn_scene_read_visible": (C.c_int, [vp, vp, C.c_uint32, c_u32_p]),
    "zn_scene_read_world": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
    "zn_scene_read_mvp": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
    "zn_scene_read_inputs": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp]),
    "zn_scene_device_buffers": (C.c_int, [vp, C.POINTER(SceneBuffers)]),
    "zn_scene_build_draw": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, vp]),
    "zn_scene_save": (C.c_int, [vp, C.c_char_p]),
    "zn_scene_load": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "zn_mesh_save": (C.c_int, [vp, C.c_char_p]),
    "zn_mesh_load": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "zn_scene_frame_host": (C.c_int, [vp, C.POINTER(FrameHostIO)]),
    "zn_scene_frame_host_submit": (C.c_int, [vp, C.POINTER(FrameHostIO)]),
    "zn_scene_frame_host_wait": (C.c_int, [vp, C.POINTER(FrameHostIO)]),
    "zn_scene_algorithmic_bytes": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_uint64)]),
    "zn_scene_launches_per_update": (C.c_int, [vp, c_u32_p]),
    "zn_scene_set_profiling": (C.c_int, [vp, C.c_int]),
    "zn_scene_level_times": (C.c_int, [vp, vp, C.c_uint32, c_float_p]),
    "zn_shard_scenes": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, c_u32_p, c_u32_p]),
    "zn_scene_index_bases": (C.c_int, [vp, C.c_uint32, vp]),
    "zn_shard_scene_by_roots": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp]),
    "zn_exchange_create": (C.c_int, [vp, C.POINTER(ExchangeDesc), C.POINTER(vp)]),
    "zn_exchange_destroy": (C.c_int, [vp, vp, C.c_uint32]),
    "zn_exchange_unbind": (C.c_int, [vp, vp, C.c_uint32]),
    "zn_exchange_bind": (C.c_int, [vp, vp, C.c_uint32, C.c_uint64]),
    "zn_exchange_expand": (C.c_int, [vp, vp, C.c_int64, vp]),
    "zn_exchange_signal": (C.c_int, [vp, vp]),
    "zn_exchange_join": (C.c_int, [vp]),
    "zn_exchange_result": (C.c_int, [vp, vp, C.POINTER(vp), c_u32_p]),
    "zn_exchange_info": (C.c_int, [vp, c_u32_p, c_u32_p, C.POINTER(C.c_uint64)]),
    "zn_mesh_create": (C.c_int, [vp, C.POINTER(MeshDesc), C.POINTER(vp)]),
    "zn_mesh_destroy": (C.c_int, [vp]),
    "zn_mesh_set_vertices": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_size_t]),
    "zn_mesh_set_models": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
    "zn_mesh_bind_scene_models": (C.c_int, [vp, vp, C.c_uint32]),
    "zn_mesh_transform": (C.c_int, [vp]),
    "zn_mesh_read": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp]),
    "zn_mesh_read_inputs": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp]),
    "zn_mesh_device_buffers": (C.c_int, [vp, C.POINTER(MeshBuffers)]),
    "zn_scene_fill_synthetic": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_float]),
    "zn_mesh_fill_synthetic": (C.c_int, [vp, C.c_uint32]),
}

_lib = None


class ZenythError(RuntimeError):
    """Raised for any non-zero status, mirroring the reference's std::runtime_error convention
    (Core/src/Application.cpp:28-30)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[zn error {code}] {message}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m zenyth_b200.build` (nvcc, sm_100a). "
            "zenyth_b200 has no CPU or Python fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError here = the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.zn_abi_version() != 2:
        raise ImportError(f"libzenyth_b200.so ABI version {lib.zn_abi_version()} != 2")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != ZN_OK:
        raise ZenythError(code, load().zn_last_error().decode("utf-8", "replace"))
