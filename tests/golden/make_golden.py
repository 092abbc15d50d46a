"""Generates tests/golden/*.npz — committed golden vectors for the scene path.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The vectors are produced by the oracle built on the REFERENCE'S OWN zenyth::math::vec3/vec4/mat4
(oracle/_ref: /root/reference/Core/src/math/{vector,matrix}.cpp compiled in place), so they carry
the reference's real SSE arithmetic to the GPU box, where /root/reference does not exist.  The
reference itself holds no golden vectors for this path, and cannot produce any: its
mat4 is stubbed.  tests/test_golden.py checks the restated oracle (CPU) and the CUDA path (GPU)
against these files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.binding import BASE_SEED, Oracle, build, camera, h8_geo_level_sizes  # noqa: E402


def main():
    build(ref=True)
    o = Oracle(reference_types=True)
    assert o.lib.zo_backend() == 1

    # 1. hierarchy: 3 levels, fan-out 8, 7 roots = 511 entities, three cameras
    sizes = h8_geo_level_sizes(7, 3)
    sc = o.synth_scene(BASE_SEED + 3, sizes, fanout=8, center_range=0.25)
    out = {k: sc[k] for k in ("position", "rotation", "scale", "aabb_center", "aabb_half", "parent", "level_offsets")}
    for name, eye in (("inside", (0, 0, 0)), ("mid", (0, 0, 1500)), ("wide", (0, 0, 3000))):
        view, proj = camera(o, eye=eye)
        r = o.scene_update(sc, view, proj)
        out[f"{name}_view"], out[f"{name}_proj"] = view, proj
        out[f"{name}_visible"] = r["visible"]
        out[f"{name}_mvp"] = r["mvp"]
    out["world"] = r["world"]
    np.savez_compressed(os.path.join(HERE, "scene_h3.npz"), **out)

    # 2. flat, ragged size (not a multiple of the 256-entity tile), random AABB centres
    fl = o.synth_scene(BASE_SEED + 2, [300], fanout=1, center_range=0.5)
    view, proj = camera(o, eye=(0, 0, 1500))
    r = o.scene_update(fl, view, proj)
    np.savez_compressed(os.path.join(HERE, "scene_flat300.npz"), view=view, proj=proj, world=r["world"], mvp=r["mvp"], visible=r["visible"],
                        **{k: fl[k] for k in ("position", "rotation", "scale", "aabb_center", "aabb_half")})

    # 3. mesh: ragged instance ranges, models = world matrices of the flat scene
    counts = np.array([40, 1, 1024, 7, 260], np.uint32)
    first = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    pos, nrm = o.synth_mesh(BASE_SEED + 4, int(first[-1]))
    models = r["world"][: counts.size].copy()
    p, n = o.mesh_transform(first, models, pos, nrm)
    np.savez_compressed(os.path.join(HERE, "mesh_ragged.npz"), first=first, models=models, pos=pos, nrm=nrm, pos_out=p, nrm_out=n)

    # 4. camera / mat4 known answers as the reference-types build computes them
    np.savez_compressed(os.path.join(HERE, "mat4_kat.npz"),
                        perspective=o.perspective(np.float32(np.pi / 2), 1.0, 1.0, 2.0),
                        look_at=o.look_at((3, -2, 5), (0, 0, 0), (0, 1, 0)),
                        euler=o.from_euler(0.3, -0.7, 1.1),
                        inverse=o.inverse(o.local_matrix((1, 2, 3), (0.3, -0.7, 1.1), (0.5, 2, 1.5))),
                        planes=o.frustum_planes(o.mat_mul(o.perspective(1.0, 1.6, 0.1, 900.0), o.look_at((3, -2, 5), (0, 0, 0), (0, 1, 0)))))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
