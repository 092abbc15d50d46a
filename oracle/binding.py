"""ORACLE — test infrastructure, not product code.

ctypes loader for the float32 oracle (oracle/libzo_oracle.so, or oracle/_ref/libzo_oracle_ref.so
built on the reference's own zenyth::math types) and the float64 shadow.  Imported only by
tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference arm); the
product package zenyth_b200 never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_RESTATED = os.path.join(HERE, "libzo_oracle.so")
LIB_REFERENCE = os.path.join(HERE, "_ref", "libzo_oracle_ref.so")
NO_PARENT = 0xFFFFFFFF
BASE_SEED = 0x5A454E59

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    if ref and os.path.isdir("/root/reference/Core/src/math"):
        subprocess.run(["make", "-s", "-C", HERE, "ref", "mat4", "loop"], check=True)


def _ptr(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One loaded oracle library. `reference_types=True` loads the oracle/_ref build."""

    def __init__(self, reference_types: bool = False):
        path = LIB_REFERENCE if reference_types else LIB_RESTATED
        if not os.path.exists(path):
            if reference_types:
                raise FileNotFoundError(path)
            build(ref=False)
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.zo_backend.restype = C.c_int
        L.zo_hardware_threads.restype = C.c_uint
        for name in ("zo_vec3_dot", "zo_vec4_dot"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [_f32p, _f32p]
        for name in ("zo_vec3_length", "zo_vec4_length", "zo_determinant"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [_f32p]
        L.zo_synth_hash.restype = C.c_uint32
        L.zo_synth_hash.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.zo_aabb_visible.restype = C.c_int
        assert L.zo_backend() == (1 if reference_types else 0)

    # -- helpers ----------------------------------------------------------------------------
    @staticmethod
    def _f(a, n=None):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
        if n is not None:
            assert a.size == n, (a.size, n)
        return a

    def hardware_threads(self) -> int:
        return int(self.lib.zo_hardware_threads())

    # -- vectors ----------------------------------------------------------------------------
    def vec3_op(self, op, a, b):
        out = np.zeros(3, np.float32)
        self.lib.zo_vec3_binary(C.c_int(op), _ptr(self._f(a, 3), np.float32), _ptr(self._f(b), np.float32), _ptr(out, np.float32))
        return out

    def vec4_op(self, op, a, b):
        out = np.zeros(4, np.float32)
        self.lib.zo_vec4_binary(C.c_int(op), _ptr(self._f(a, 4), np.float32), _ptr(self._f(b), np.float32), _ptr(out, np.float32))
        return out

    def vec3_dot(self, a, b):
        return np.float32(self.lib.zo_vec3_dot(self._f(a, 3), self._f(b, 3)))

    def vec4_dot(self, a, b):
        return np.float32(self.lib.zo_vec4_dot(self._f(a, 4), self._f(b, 4)))

    def vec3_length(self, a):
        return np.float32(self.lib.zo_vec3_length(self._f(a, 3)))

    def vec4_length(self, a):
        return np.float32(self.lib.zo_vec4_length(self._f(a, 4)))

    # -- mat4 contract ----------------------------------------------------------------------
    def sincos(self, x):
        x = np.ascontiguousarray(np.asarray(x, np.float32).reshape(-1))
        s = np.empty_like(x)
        c = np.empty_like(x)
        self.lib.zo_sincos_array(C.c_uint32(x.size), _ptr(x, np.float32), _ptr(s, np.float32), _ptr(c, np.float32))
        return s, c

    def _m(self, fn, *args):
        out = np.zeros(16, np.float32)
        getattr(self.lib, fn)(*args, _ptr(out, np.float32))
        return out

    def _v(self, a, n):
        return _ptr(self._f(a, n), np.float32)

    def identity(self):
        return self._m("zo_identity")

    def diagonal(self, d):
        return self._m("zo_diagonal", C.c_float(d))

    def from_basis(self, r, u, f):
        return self._m("zo_from_basis", self._v(r, 3), self._v(u, 3), self._v(f, 3))

    def from_axis_angle(self, axis, angle):
        return self._m("zo_from_axis_angle", self._v(axis, 3), C.c_float(angle))

    def from_euler(self, pitch, yaw, roll):
        return self._m("zo_from_euler", C.c_float(pitch), C.c_float(yaw), C.c_float(roll))

    def from_scale(self, s):
        return self._m("zo_from_scale", self._v(s, 3))

    def from_translation(self, t):
        return self._m("zo_from_translation", self._v(t, 3))

    def mat_add(self, a, b):
        return self._m("zo_mat_add", self._v(a, 16), self._v(b, 16))

    def mat_sub(self, a, b):
        return self._m("zo_mat_sub", self._v(a, 16), self._v(b, 16))

    def mat_scale(self, a, s):
        return self._m("zo_mat_scale", self._v(a, 16), C.c_float(s))

    def mat_div(self, a, s):
        return self._m("zo_mat_div", self._v(a, 16), C.c_float(s))

    def mat_mul(self, a, b):
        return self._m("zo_mat_mul", self._v(a, 16), self._v(b, 16))

    def mat_mul_vec4(self, a, v):
        out = np.zeros(4, np.float32)
        self.lib.zo_mat_mul_vec4(self._v(a, 16), self._v(v, 4), _ptr(out, np.float32))
        return out

    def transpose(self, a):
        return self._m("zo_transpose", self._v(a, 16))

    def determinant(self, a):
        return np.float32(self.lib.zo_determinant(self._f(a, 16)))

    def inverse(self, a):
        return self._m("zo_inverse", self._v(a, 16))

    def inverse_transpose(self, a):
        return self._m("zo_inverse_transpose", self._v(a, 16))

    def transform_point(self, a, p):
        out = np.zeros(3, np.float32)
        self.lib.zo_transform_point(self._v(a, 16), self._v(p, 3), _ptr(out, np.float32))
        return out

    def transform_normal(self, a, n):
        out = np.zeros(3, np.float32)
        self.lib.zo_transform_normal(self._v(a, 16), self._v(n, 3), _ptr(out, np.float32))
        return out

    def look_at(self, eye, center, up):
        return self._m("zo_look_at", self._v(eye, 3), self._v(center, 3), self._v(up, 3))

    def perspective(self, fov_y, aspect, near_z, far_z):
        return self._m("zo_perspective", C.c_float(fov_y), C.c_float(aspect), C.c_float(near_z), C.c_float(far_z))

    def orthographic(self, l, r, b, t, n, f):
        return self._m("zo_orthographic", *(C.c_float(v) for v in (l, r, b, t, n, f)))

    def perspective_reverse_z(self, fov_y, aspect, near_z):
        return self._m("zo_perspective_reverse_z", C.c_float(fov_y), C.c_float(aspect), C.c_float(near_z))

    # -- scene pieces -----------------------------------------------------------------------
    def local_matrix(self, t, r, s):
        return self._m("zo_local_matrix", self._v(t, 3), self._v(r, 3), self._v(s, 3))

    def frustum_planes(self, view_proj):
        out = np.zeros(24, np.float32)
        self.lib.zo_frustum_planes(self._v(view_proj, 16), _ptr(out, np.float32))
        return out.reshape(6, 4)

    def world_aabb(self, m, c, e):
        co = np.zeros(3, np.float32)
        eo = np.zeros(3, np.float32)
        self.lib.zo_world_aabb(self._v(m, 16), self._v(c, 3), self._v(e, 3), _ptr(co, np.float32), _ptr(eo, np.float32))
        return co, eo

    def aabb_visible(self, view_proj, c, e) -> bool:
        return bool(self.lib.zo_aabb_visible(self._v(view_proj, 16), self._v(c, 3), self._v(e, 3)))

    def scene_update(self, scene: dict, view, proj, threads: int = 1, want_mvp: bool = True):
        """scene: dict with position/rotation/scale/aabb_center/aabb_half [E,3] f32, parent [E] u32
        (or None), level_offsets [L+1] u32 (or None).  Returns dict(world, mvp, flags, visible)."""
        n = int(scene["position"].shape[0])
        world = np.empty((n, 16), np.float32)
        mvp = np.empty((n, 16), np.float32) if want_mvp else None
        flags = np.empty(n, np.uint8)
        visible = np.empty(n, np.uint32)
        count = C.c_uint32(0)
        lo = scene.get("level_offsets")
        par = scene.get("parent")
        self.lib.zo_scene_update(
            C.c_uint32(n), _ptr(lo, np.uint32), C.c_uint32(0 if lo is None else lo.size - 1),
            _ptr(scene["position"], np.float32), _ptr(scene["rotation"], np.float32), _ptr(scene["scale"], np.float32),
            _ptr(scene["aabb_center"], np.float32), _ptr(scene["aabb_half"], np.float32), _ptr(par, np.uint32),
            self._v(view, 16), self._v(proj, 16),
            _ptr(world, np.float32), _ptr(mvp, np.float32), _ptr(flags, np.uint8),
            _ptr(visible, np.uint32), C.byref(count), C.c_uint(threads))
        return {"world": world, "mvp": mvp, "flags": flags, "visible": visible[: count.value].copy()}

    def mesh_transform(self, first_vertex, model, pos, nrm, threads: int = 1):
        pos_out = np.empty_like(pos)
        nrm_out = np.empty_like(nrm)
        self.lib.zo_mesh_transform(C.c_uint32(first_vertex.size - 1), _ptr(first_vertex, np.uint32),
                                   _ptr(np.ascontiguousarray(model, np.float32), np.float32),
                                   _ptr(pos, np.float32), _ptr(nrm, np.float32),
                                   _ptr(pos_out, np.float32), _ptr(nrm_out, np.float32), C.c_uint(threads))
        return pos_out, nrm_out

    # -- synthetic inputs -------------------------------------------------------------------
    def synth_hash(self, seed, index, field) -> int:
        return int(self.lib.zo_synth_hash(seed & 0xFFFFFFFF, index, field))

    def synth_scene(self, seed: int, level_sizes, fanout: int = 8, center_range: float = 0.0, threads: int = 0) -> dict:
        level_offsets = np.concatenate([[0], np.cumsum(np.asarray(level_sizes, np.uint64))]).astype(np.uint32)
        n = int(level_offsets[-1])
        s = {k: np.empty((n, 3), np.float32) for k in ("position", "rotation", "scale", "aabb_center", "aabb_half")}
        s["parent"] = np.empty(n, np.uint32)
        s["level_offsets"] = level_offsets
        self.lib.zo_synth_scene(C.c_uint32(seed & 0xFFFFFFFF), _ptr(level_offsets, np.uint32), C.c_uint32(level_offsets.size - 1),
                                C.c_uint32(fanout), C.c_float(center_range),
                                _ptr(s["position"], np.float32), _ptr(s["rotation"], np.float32), _ptr(s["scale"], np.float32),
                                _ptr(s["aabb_center"], np.float32), _ptr(s["aabb_half"], np.float32), _ptr(s["parent"], np.uint32),
                                C.c_uint(threads or self.hardware_threads()))
        return s

    def synth_mesh(self, seed: int, vertex_count: int, threads: int = 0):
        pos = np.empty((vertex_count, 3), np.float32)
        nrm = np.empty((vertex_count, 3), np.float32)
        self.lib.zo_synth_mesh(C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(vertex_count), _ptr(pos, np.float32), _ptr(nrm, np.float32),
                               C.c_uint(threads or self.hardware_threads()))
        return pos, nrm

    # -- float64 shadow ---------------------------------------------------------------------
    def sincos64(self, x):
        x64 = np.asarray(x, np.float32).astype(np.float64)
        return np.sin(x64), np.cos(x64)

    def from_euler64(self, pitch, yaw, roll):
        out = np.zeros(16, np.float64)
        self.lib.zo64_from_euler(C.c_float(pitch), C.c_float(yaw), C.c_float(roll), _ptr(out, np.float64))
        return out

    def local_matrix64(self, t, r, s):
        out = np.zeros(16, np.float64)
        self.lib.zo64_local_matrix(self._v(t, 3), self._v(r, 3), self._v(s, 3), _ptr(out, np.float64))
        return out

    def scene_update64(self, scene: dict, view, proj, eps: float = 1e-5, want_matrices: bool = True):
        n = int(scene["position"].shape[0])
        world = np.empty((n, 16), np.float64) if want_matrices else None
        mvp = np.empty((n, 16), np.float64) if want_matrices else None
        vis = np.empty(n, np.uint8)
        bnd = np.empty(n, np.uint8)
        self.lib.zo64_scene_update(
            C.c_uint32(n), _ptr(scene["position"], np.float32), _ptr(scene["rotation"], np.float32), _ptr(scene["scale"], np.float32),
            _ptr(scene["aabb_center"], np.float32), _ptr(scene["aabb_half"], np.float32), _ptr(scene.get("parent"), np.uint32),
            self._v(view, 16), self._v(proj, 16), C.c_double(eps),
            _ptr(world, np.float64), _ptr(mvp, np.float64), _ptr(vis, np.uint8), _ptr(bnd, np.uint8))
        return {"world": world, "mvp": mvp, "visible": vis, "boundary": bnd}

    def mesh_transform64(self, first_vertex, model, pos, nrm):
        pos_out = np.empty(pos.shape, np.float64)
        nrm_out = np.empty(nrm.shape, np.float64)
        self.lib.zo64_mesh_transform(C.c_uint32(first_vertex.size - 1), _ptr(first_vertex, np.uint32),
                                     _ptr(np.ascontiguousarray(model, np.float32), np.float32),
                                     _ptr(pos, np.float32), _ptr(nrm, np.float32),
                                     _ptr(pos_out, np.float64), _ptr(nrm_out, np.float64))
        return pos_out, nrm_out


def h8_geo_level_sizes(roots: int = 7, levels: int = 8, fanout: int = 8):
    """SURVEY.md §8(d) config 3/5 shape: level l holds roots * fanout**l entities."""
    return [roots * fanout ** l for l in range(levels)]


def camera(oracle: Oracle, eye=(0.0, 0.0, 1500.0), center=None, fov_deg=60.0, aspect=1280.0 / 720.0, near_z=0.1, far_z=5000.0):
    """The 'mid' camera of SURVEY.md §8(d): looks toward -Z from `eye`."""
    eye = np.asarray(eye, np.float32)
    center = np.asarray(center, np.float32) if center is not None else eye + np.asarray([0, 0, -1], np.float32)
    view = oracle.look_at(eye, center, (0.0, 1.0, 0.0))
    proj = oracle.perspective(np.float32(fov_deg) * np.float32(np.pi) / np.float32(180.0), aspect, near_z, far_z)
    return view, proj
