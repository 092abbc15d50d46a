"""Host-side mirror of the engine-idiom API (include/zenyth_b200.hpp) over the C ABI.

Class and method names follow the reference's engine layer (namespace Zenyth, PascalCase methods,
Desc aggregates — Core/include/Application.hpp:9-14,26-40); the maths keeps the reference's
snake_case (Core/include/math/matrix.hpp:19-55).  Everything here is plumbing: numpy arrays in,
ctypes calls to libzenyth_b200.so, numpy arrays out.  No arithmetic of the path happens in Python.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import ZN_NO_PARENT, ZenythError, check

__all__ = ["GpuContext", "Transform", "CameraDesc", "Camera", "Scene", "Mesh", "mat4", "SortHierarchy", "ZN_NO_PARENT", "ZenythError"]


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class mat4:
    """The camera-side completion of zenyth::math::mat4 (static factories, matrix.hpp:32,52-55).
    Matrices are float32[16], column-major, as mat4::data()."""

    @staticmethod
    def look_at(eye, center, up) -> np.ndarray:
        out = np.zeros(16, np.float32)
        _capi.load().zn_mat4_look_at(_p(_f32(eye, 3)), _p(_f32(center, 3)), _p(_f32(up, 3)), _p(out))
        return out

    @staticmethod
    def perspective(fov_y: float, aspect: float, near_z: float, far_z: float) -> np.ndarray:
        out = np.zeros(16, np.float32)
        _capi.load().zn_mat4_perspective(fov_y, aspect, near_z, far_z, _p(out))
        return out

    @staticmethod
    def orthographic(left, right, bottom, top, near_z, far_z) -> np.ndarray:
        out = np.zeros(16, np.float32)
        _capi.load().zn_mat4_orthographic(left, right, bottom, top, near_z, far_z, _p(out))
        return out

    @staticmethod
    def perspective_reverse_z(fov_y: float, aspect: float, near_z: float) -> np.ndarray:
        out = np.zeros(16, np.float32)
        _capi.load().zn_mat4_perspective_reverse_z(fov_y, aspect, near_z, _p(out))
        return out

    @staticmethod
    def mul(a, b) -> np.ndarray:
        out = np.zeros(16, np.float32)
        _capi.load().zn_mat4_mul(_p(_f32(a, 16)), _p(_f32(b, 16)), _p(out))
        return out

    @staticmethod
    def frustum_planes(view_proj) -> np.ndarray:
        out = np.zeros(24, np.float32)
        _capi.load().zn_frustum_planes(_p(_f32(view_proj, 16)), _p(out))
        return out.reshape(6, 4)


def SortHierarchy(parent):
    """Entities in creation order (parent[e] = index of e's parent, ZN_NO_PARENT for roots) ->
    (order, new_parent, level_offsets): the level-ordered, sorted-by-parent layout the Scene calls
    require (zn_hierarchy_sort; host only, no GPU).  Per-entity arrays go along as array[order]."""
    lib = _capi.load()
    parent = np.ascontiguousarray(parent, np.uint32)
    n = int(parent.size)
    order = np.empty(n, np.uint32)
    new_parent = np.empty(n, np.uint32)
    levels = C.c_uint32(0)
    capacity = 66
    while True:
        offsets = np.empty(capacity, np.uint32)
        rc = lib.zn_hierarchy_sort(_p(parent), n, _p(order), _p(new_parent), _p(offsets), capacity, C.byref(levels))
        if rc != 0 and levels.value + 1 > capacity:      # deeper than guessed: the count came back, retry once
            capacity = levels.value + 1
            continue
        check(rc)
        return order, new_parent, offsets[: levels.value + 1].copy()


class GpuContext:
    """One per GPU (zn_ctx).  Move-only in the C++ wrapper; here: explicit Close() or context manager."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _capi.load()
        h = C.c_void_p()
        check(self._lib.zn_ctx_create(device, C.byref(h)))
        self._h = h
        self.device = device
        if stream is not None:
            self.SetStream(stream)

    def SetStream(self, cuda_stream: int | None) -> None:
        check(self._lib.zn_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def GetStream(self) -> int:
        """The cudaStream_t (as an integer) all work of this context is enqueued on."""
        p = C.c_void_p()
        check(self._lib.zn_ctx_get_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def Sync(self) -> None:
        check(self._lib.zn_ctx_sync(self._h))

    def SmCount(self) -> int:
        n = C.c_int()
        check(self._lib.zn_ctx_sm_count(self._h, C.byref(n)))
        return n.value

    def HostAlloc(self, shape, dtype) -> np.ndarray:
        """Page-locked numpy array (freed with the context)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        check(self._lib.zn_host_alloc(self._h, n, C.byref(ptr)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_uint8 * max(n, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # -- peer memory (multi-GPU, one process per GPU) -------------------------------------------
    def DeviceAlloc(self, nbytes: int) -> int:
        """A zero-filled device allocation other ranks can map (zn_device_alloc); returns the pointer."""
        ptr = C.c_void_p()
        check(self._lib.zn_device_alloc(self._h, nbytes, C.byref(ptr)))
        return int(ptr.value)

    def DeviceFree(self, ptr: int) -> None:
        check(self._lib.zn_device_free(self._h, C.c_void_p(ptr)))

    def IpcExport(self, ptr: int) -> bytes:
        buf = (C.c_ubyte * 64)()
        check(self._lib.zn_ipc_export(self._h, C.c_void_p(ptr), buf))
        return bytes(buf)

    def IpcOpen(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        ptr = C.c_void_p()
        check(self._lib.zn_ipc_open(self._h, buf, C.byref(ptr)))
        return int(ptr.value)

    def IpcClose(self, ptr: int) -> None:
        check(self._lib.zn_ipc_close(self._h, C.c_void_p(ptr)))

    def Close(self) -> None:
        if getattr(self, "_h", None):
            for ptr in getattr(self, "_pinned", []):
                self._lib.zn_host_free(self._h, ptr)
            self._pinned = []
            self._lib.zn_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.Close()

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass


@dataclass
class Transform:
    """Local TRS of one entity.  rotation = (pitch, yaw, roll) radians, as mat4::from_euler."""
    position: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = (0.0, 0.0, 0.0)
    scale: tuple = (1.0, 1.0, 1.0)


@dataclass
class CameraDesc:
    eye: tuple = (0.0, 0.0, 0.0)
    center: tuple = (0.0, 0.0, -1.0)
    up: tuple = (0.0, 1.0, 0.0)
    fov_y_rad: float = float(np.float32(60.0) * np.float32(np.pi) / np.float32(180.0))
    aspect: float = 1280.0 / 720.0   # Sandbox/include/SandboxApp.hpp:6
    near_z: float = 0.1
    far_z: float = 5000.0


class Camera:
    def __init__(self, desc: CameraDesc | None = None, view=None, proj=None):
        desc = desc or CameraDesc()
        self.desc = desc
        self.view = _f32(view, 16) if view is not None else mat4.look_at(desc.eye, desc.center, desc.up)
        self.proj = _f32(proj, 16) if proj is not None else mat4.perspective(desc.fov_y_rad, desc.aspect, desc.near_z, desc.far_z)

    def ViewProj(self) -> np.ndarray:
        return mat4.mul(self.proj, self.view)

    def Resize(self, width: int, height: int) -> None:
        """What IRenderer::Resize (Core/include/IRenderer.hpp:11) feeds: a new aspect ratio."""
        self.desc.aspect = float(width) / float(height)
        self.proj = mat4.perspective(self.desc.fov_y_rad, self.desc.aspect, self.desc.near_z, self.desc.far_z)


class Scene:
    """Level-ordered entity hierarchy resident on one GPU (zn_scene)."""

    def __init__(self, ctx: GpuContext, entity_count: int, level_offsets=None):
        self._lib = _capi.load()
        self.ctx = ctx
        self.entity_count = int(entity_count)
        lo = None if level_offsets is None else np.ascontiguousarray(level_offsets, np.uint32)
        self.level_offsets = lo
        desc = _capi.SceneDesc(self.entity_count, 0 if lo is None else lo.size - 1, _p(lo))
        h = C.c_void_p()
        check(self._lib.zn_scene_create(ctx._h, C.byref(desc), C.byref(h)))
        self._h = h

    # -- uploads ------------------------------------------------------------------------------
    @staticmethod
    def _records(a, count):
        """Accept [count,3] packed or [count,4] (16-byte vec3) float32 arrays."""
        if a is None:
            return None, 0
        a = np.ascontiguousarray(a, np.float32)
        assert a.ndim == 2 and a.shape[0] == count and a.shape[1] in (3, 4), a.shape
        return a, a.shape[1] * 4

    def SetTransforms(self, position=None, rotation=None, scale=None, first: int = 0, count: int | None = None) -> None:
        count = self._count(count, position, rotation, scale)
        arrs = [self._records(a, count) for a in (position, rotation, scale)]
        strides = {s for _, s in arrs if s}
        assert len(strides) <= 1, "position/rotation/scale must share one stride"
        stride = strides.pop() if strides else 12
        check(self._lib.zn_scene_set_transforms(self._h, first, count, _p(arrs[0][0]), _p(arrs[1][0]), _p(arrs[2][0]), stride))

    def SetTransform(self, index: int, t: Transform) -> None:
        self.SetTransforms(np.asarray([t.position], np.float32), np.asarray([t.rotation], np.float32),
                           np.asarray([t.scale], np.float32), first=index, count=1)

    def SetBounds(self, aabb_center=None, aabb_half=None, first: int = 0, count: int | None = None) -> None:
        count = self._count(count, aabb_center, aabb_half)
        arrs = [self._records(a, count) for a in (aabb_center, aabb_half)]
        strides = {s for _, s in arrs if s}
        stride = strides.pop() if strides else 12
        check(self._lib.zn_scene_set_bounds(self._h, first, count, _p(arrs[0][0]), _p(arrs[1][0]), stride))

    def SetParents(self, parent, first: int = 0) -> None:
        parent = np.ascontiguousarray(parent, np.uint32)
        check(self._lib.zn_scene_set_parents(self._h, first, parent.size, _p(parent)))

    def SetCamera(self, camera: Camera) -> None:
        check(self._lib.zn_scene_set_camera(self._h, _p(camera.view), _p(camera.proj)))

    def BindVisibleOutput(self, visible_ptr: int | None, capacity: int = 0, count_ptr: int | None = None) -> None:
        """Redirect the visible list / count into caller-owned device memory (raw device pointers)."""
        check(self._lib.zn_scene_bind_visible_output(self._h, C.c_void_p(visible_ptr or 0), capacity, C.c_void_p(count_ptr or 0)))

    def BindVisibilityRecord(self, record_ptr: int | None) -> None:
        """Subsequent updates write the frame's visibility record into caller-owned device memory."""
        check(self._lib.zn_scene_bind_visibility_record(self._h, C.c_void_p(record_ptr or 0)))

    def BindPeerRecords(self, peer_ptrs) -> None:
        """Every update also stores the visibility record into these mapped PEER-GPU buffers."""
        peer_ptrs = list(peer_ptrs)
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*peer_ptrs)
        check(self._lib.zn_scene_bind_peer_records(self._h, arr, len(peer_ptrs)))

    def ExpandVisible(self, records_ptr: int, record_count: int, record_stride_words: int, lists_ptr: int, list_stride: int,
                      counts_ptr: int, index_bases=None, skip_index: int = -1) -> None:
        """Expand all-gathered visibility records into ordered index lists (zn_scene_expand_visible)."""
        bases = None if index_bases is None else np.ascontiguousarray(index_bases, np.uint32)
        check(self._lib.zn_scene_expand_visible(self._h, C.c_void_p(records_ptr), record_count, record_stride_words, skip_index,
                                                _p(bases), C.c_void_p(lists_ptr), list_stride, C.c_void_p(counts_ptr)))

    def BindPeerSignal(self, flag_ptrs) -> None:
        """The flag words (one per rank, mapped peer memory) that the NEXT ExpandVisibleSignalled / WaitFlags launch
        raises to the signal value — the update itself raises nothing (zenyth_b200.h), so the last frame of a run
        must be followed by one more such call."""
        flag_ptrs = list(flag_ptrs)
        arr = (C.c_void_p * max(len(flag_ptrs), 1))(*flag_ptrs)
        check(self._lib.zn_scene_bind_peer_signal(self._h, arr, len(flag_ptrs)))

    def WaitFlags(self, flags_ptr: int, flag_stride_words: int, flag_count: int, expected: int, status_ptr: int, timeout_ms: int = 2000) -> None:
        """A one-CTA kernel on the context stream that waits until every flag has reached `expected` (bounded)."""
        check(self._lib.zn_scene_wait_flags(self._h, C.c_void_p(flags_ptr), flag_stride_words, flag_count, expected & 0xFFFFFFFF,
                                            timeout_ms, C.c_void_p(status_ptr)))

    def SetSignalValue(self, value: int) -> None:
        check(self._lib.zn_scene_set_signal_value(self._h, value & 0xFFFFFFFF))

    def ExpandVisibleSignalled(self, records_ptr: int, record_count: int, record_stride_words: int, lists_ptr: int, list_stride: int,
                               counts_ptr: int, flags_ptr: int, flag_stride_words: int, expected: int, status_ptr: int,
                               index_bases=None, skip_index: int = -1, skip_count: int = 1, records_per_flag: int = 1,
                               timeout_ms: int = 2000) -> None:
        """ExpandVisible whose CTAs first wait for the owners' flags (zn_scene_expand_visible_signalled)."""
        bases = None if index_bases is None else np.ascontiguousarray(index_bases, np.uint32)
        check(self._lib.zn_scene_expand_visible_signalled(
            self._h, C.c_void_p(records_ptr), record_count, record_stride_words, skip_index, skip_count, _p(bases), C.c_void_p(lists_ptr),
            list_stride, C.c_void_p(counts_ptr), C.c_void_p(flags_ptr), flag_stride_words, records_per_flag, expected & 0xFFFFFFFF,
            timeout_ms, C.c_void_p(status_ptr)))

    def SetExpandMode(self, worklist: bool) -> None:
        """How records are expanded with this scene: one launch with a warp per group (default; dense visible sets)
        or classification + work list (sparse visible sets).  zn_scene_set_expand_mode."""
        check(self._lib.zn_scene_set_expand_mode(self._h, 1 if worklist else 0))

    def SetLevelFusion(self, on: bool) -> None:
        """The tiny leading levels of a hierarchy in one launch that keeps their matrices on chip (default on; results
        are bit-identical either way).  zn_scene_set_level_fusion."""
        check(self._lib.zn_scene_set_level_fusion(self._h, 1 if on else 0))

    def SetIndexBase(self, base: int) -> None:
        check(self._lib.zn_scene_set_index_base(self._h, base))

    def Load(self, arrays: dict) -> None:
        """arrays: position/rotation/scale/aabb_center/aabb_half [E,3], parent [E] (optional)."""
        self.SetTransforms(arrays["position"], arrays["rotation"], arrays["scale"])
        self.SetBounds(arrays["aabb_center"], arrays["aabb_half"])
        if arrays.get("parent") is not None:
            self.SetParents(arrays["parent"])

    def FillSynthetic(self, seed: int, fanout: int = 8, aabb_center_range: float = 0.0) -> None:
        check(self._lib.zn_scene_fill_synthetic(self._h, seed & 0xFFFFFFFF, fanout, aabb_center_range))

    def _count(self, count, *arrs):
        if count is not None:
            return int(count)
        for a in arrs:
            if a is not None:
                return int(np.asarray(a).shape[0])
        return 0

    # -- frame --------------------------------------------------------------------------------
    def Update(self) -> None:
        """One frame, asynchronous: world, MVP, visible list (zn_scene_update).  Without a camera: world only."""
        check(self._lib.zn_scene_update(self._h))

    def UpdateWorld(self) -> None:
        """TRS -> world over the hierarchy only (zn_scene_update_world): what OnUpdate(dt) calls."""
        check(self._lib.zn_scene_update_world(self._h))

    def Cull(self, camera: "Camera", mvp_ptr: int | None = None, visible_ptr: int | None = None, capacity: int = 0,
             count_ptr: int | None = None) -> None:
        """One camera pass over the existing world matrices (zn_scene_cull): MVP, cull, ordered list.
        Optional raw device pointers redirect this pass's results; default: the scene's own buffers."""
        out = None
        if mvp_ptr or visible_ptr or count_ptr:
            out = C.byref(_capi.CullOutput(C.c_void_p(mvp_ptr or 0), C.c_void_p(visible_ptr or 0), capacity, C.c_void_p(count_ptr or 0)))
        check(self._lib.zn_scene_cull(self._h, _p(camera.view), _p(camera.proj), out))

    def ClearCamera(self) -> None:
        check(self._lib.zn_scene_clear_camera(self._h))

    def SetTicketOrigin(self, origin: int) -> None:
        """Test support: restart the kernels' 32-bit tile-ticket counters at `origin` (zn_scene_set_ticket_origin)."""
        check(self._lib.zn_scene_set_ticket_origin(self._h, origin & 0xFFFFFFFF))

    def CullTimesMs(self):
        """(cull kernel ms, emit kernel ms) of the last Cull; needs SetProfiling(True)."""
        a, b = C.c_float(), C.c_float()
        check(self._lib.zn_scene_cull_times(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def FrameHost(self, position=None, rotation=None, scale=None, world=None, mvp=None, visible=None, first: int = 0, count: int = 0) -> int:
        """End-to-end frame with host buffers (zn_scene_frame_host). Returns the visible count.
        The input arrays cover entities [first, first+count) (count 0: to the end); None = keep what is resident."""
        io = _capi.FrameHostIO(_p(position), _p(rotation), _p(scale), _p(world), _p(mvp), _p(visible), 0, first, count)
        check(self._lib.zn_scene_frame_host(self._h, C.byref(io)))
        return int(io.visible_count)

    def SaveFile(self, path: str) -> None:
        """Write the scene's inputs as a ZNSCENE1 file (zn_scene_save)."""
        check(self._lib.zn_scene_save(self._h, str(path).encode()))

    @classmethod
    def LoadFile(cls, ctx: "GpuContext", path: str) -> "Scene":
        """Create a scene from a ZNSCENE1 file (zn_scene_load)."""
        lib = _capi.load()
        h = C.c_void_p()
        check(lib.zn_scene_load(ctx._h, str(path).encode(), C.byref(h)))
        with open(path, "rb") as f:
            f.seek(8)
            n, levels = np.frombuffer(f.read(8), np.uint32)
            offsets = np.frombuffer(f.read(4 * (int(levels) + 1)), np.uint32).copy()
        self = cls.__new__(cls)
        self._lib, self.ctx, self._h = lib, ctx, h
        self.entity_count = int(n)
        self.level_offsets = offsets if levels > 1 else None
        return self

    def BuildDraw(self, index_count_per_instance: int, instance_mvp_ptr: int, instance_capacity: int, draw_args_ptr: int) -> None:
        """Visible list -> packed per-instance MVP buffer + indexed-indirect draw arguments (device pointers)."""
        check(self._lib.zn_scene_build_draw(self._h, index_count_per_instance, C.c_void_p(instance_mvp_ptr), instance_capacity,
                                            C.c_void_p(draw_args_ptr)))

    def SubmitFrameHost(self, position=None, rotation=None, scale=None, first: int = 0, count: int = 0) -> None:
        """Pipelined end-to-end frame (zn_scene_frame_host_submit): at most two in flight.  Arrays that
        are None are not uploaded (a frame that only spins its entities passes `rotation` alone)."""
        io = _capi.FrameHostIO(_p(position), _p(rotation), _p(scale), None, None, None, 0, first, count)
        check(self._lib.zn_scene_frame_host_submit(self._h, C.byref(io)))

    def WaitFrameHost(self, visible=None) -> int:
        """Complete the oldest submitted frame; returns its visible count (and fills `visible`)."""
        io = _capi.FrameHostIO(None, None, None, None, None, _p(visible), 0, 0, 0)
        check(self._lib.zn_scene_frame_host_wait(self._h, C.byref(io)))
        return int(io.visible_count)

    # -- results ------------------------------------------------------------------------------
    def VisibleCount(self) -> int:
        n = C.c_uint32()
        check(self._lib.zn_scene_visible_count(self._h, C.byref(n)))
        return n.value

    def Visible(self) -> np.ndarray:
        n = self.VisibleCount()
        out = np.empty(n, np.uint32)
        got = C.c_uint32()
        check(self._lib.zn_scene_read_visible(self._h, _p(out), n, C.byref(got)))
        return out[: got.value]

    def World(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.entity_count - first if count is None else count
        out = np.empty((count, 16), np.float32)
        check(self._lib.zn_scene_read_world(self._h, first, count, _p(out)))
        return out

    def Mvp(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.entity_count - first if count is None else count
        out = np.empty((count, 16), np.float32)
        check(self._lib.zn_scene_read_mvp(self._h, first, count, _p(out)))
        return out

    def Inputs(self, first: int = 0, count: int | None = None) -> dict:
        count = self.entity_count - first if count is None else count
        out = {k: np.empty((count, 3), np.float32) for k in ("position", "rotation", "scale", "aabb_center", "aabb_half")}
        out["parent"] = np.empty(count, np.uint32)
        check(self._lib.zn_scene_read_inputs(self._h, first, count, _p(out["position"]), _p(out["rotation"]), _p(out["scale"]),
                                             _p(out["aabb_center"]), _p(out["aabb_half"]), _p(out["parent"])))
        return out

    def DeviceBuffers(self) -> _capi.SceneBuffers:
        b = _capi.SceneBuffers()
        check(self._lib.zn_scene_device_buffers(self._h, C.byref(b)))
        return b

    def AlgorithmicBytes(self, visible_count: int) -> int:
        n = C.c_uint64()
        check(self._lib.zn_scene_algorithmic_bytes(self._h, visible_count, C.byref(n)))
        return n.value

    def LaunchesPerUpdate(self) -> int:
        n = C.c_uint32()
        check(self._lib.zn_scene_launches_per_update(self._h, C.byref(n)))
        return n.value

    def SetProfiling(self, enable: bool) -> None:
        check(self._lib.zn_scene_set_profiling(self._h, int(enable)))

    def LevelTimesMs(self):
        """(per-level kernel ms of the last Update, scan+emit tail ms); needs SetProfiling(True)."""
        n = 1 if self.level_offsets is None else self.level_offsets.size - 1
        out = np.zeros(n, np.float32)
        tail = C.c_float()
        check(self._lib.zn_scene_level_times(self._h, _p(out), n, C.byref(tail)))
        return out, tail.value

    def Close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.zn_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass


class Mesh:
    """A batch of instanced vertices (zn_mesh): positions + normals transformed per instance."""

    def __init__(self, ctx: GpuContext, vertex_count: int, instance_count: int, instance_first_vertex=None):
        self._lib = _capi.load()
        self.ctx = ctx
        self.vertex_count = int(vertex_count)
        self.instance_count = int(instance_count)
        fv = None if instance_first_vertex is None else np.ascontiguousarray(instance_first_vertex, np.uint32)
        desc = _capi.MeshDesc(self.vertex_count, self.instance_count, _p(fv))
        h = C.c_void_p()
        check(self._lib.zn_mesh_create(ctx._h, C.byref(desc), C.byref(h)))
        self._h = h

    def SetVertices(self, position=None, normal=None, first: int = 0) -> None:
        count = int((position if position is not None else normal).shape[0])
        p, sp = Scene._records(position, count)
        n, sn = Scene._records(normal, count)
        stride = sp or sn or 12
        check(self._lib.zn_mesh_set_vertices(self._h, first, count, _p(p), _p(n), stride))

    def SetModels(self, models, first: int = 0) -> None:
        models = _f32(models).reshape(-1, 16)
        check(self._lib.zn_mesh_set_models(self._h, first, models.shape[0], _p(models)))

    def BindSceneModels(self, scene: Scene, first_entity: int = 0) -> None:
        check(self._lib.zn_mesh_bind_scene_models(self._h, scene._h, first_entity))
        self._scene = scene   # keep alive

    def FillSynthetic(self, seed: int) -> None:
        check(self._lib.zn_mesh_fill_synthetic(self._h, seed & 0xFFFFFFFF))

    def Transform(self) -> None:
        check(self._lib.zn_mesh_transform(self._h))

    def Read(self, first: int = 0, count: int | None = None):
        count = self.vertex_count - first if count is None else count
        pos = np.empty((count, 3), np.float32)
        nrm = np.empty((count, 3), np.float32)
        check(self._lib.zn_mesh_read(self._h, first, count, _p(pos), _p(nrm)))
        return pos, nrm

    def Inputs(self, first: int = 0, count: int | None = None):
        count = self.vertex_count - first if count is None else count
        pos = np.empty((count, 3), np.float32)
        nrm = np.empty((count, 3), np.float32)
        check(self._lib.zn_mesh_read_inputs(self._h, first, count, _p(pos), _p(nrm)))
        return pos, nrm

    def SaveFile(self, path: str) -> None:
        """Write the batch (ranges, vertices, the model matrices in use) as a ZNMESH01 file (zn_mesh_save)."""
        check(self._lib.zn_mesh_save(self._h, str(path).encode()))

    @classmethod
    def LoadFile(cls, ctx: "GpuContext", path: str) -> "Mesh":
        """Create a mesh batch from a ZNMESH01 file (zn_mesh_load)."""
        lib = _capi.load()
        h = C.c_void_p()
        check(lib.zn_mesh_load(ctx._h, str(path).encode(), C.byref(h)))
        with open(path, "rb") as f:
            f.seek(8)
            v, i = np.frombuffer(f.read(8), np.uint32)
        self = cls.__new__(cls)
        self._lib, self.ctx, self._h = lib, ctx, h
        self.vertex_count, self.instance_count = int(v), int(i)
        return self

    def DeviceBuffers(self) -> _capi.MeshBuffers:
        b = _capi.MeshBuffers()
        check(self._lib.zn_mesh_device_buffers(self._h, C.byref(b)))
        return b

    def Close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.zn_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass
