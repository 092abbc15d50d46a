"""One-GPU probe: mesh batches made of many SMALL instances (tiles that span several instances).
M vertices/s and the fraction of the measured HBM peak at 48 algorithmic bytes per vertex (+ 64 per
instance for its model matrix).  Not a bench line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zenyth_b200 as zn  # noqa: E402

peak = 6454.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = zn.GpuContext(0, stream=stream.cuda_stream)
out = {}
for verts, inst, table in ((1024, 65536, False), (36, 1 << 20, False), (8, 1 << 22, False), (4, 1 << 23, False), (3, 1 << 23, False),
                           (36, 1 << 20, True), (8, 1 << 22, True)):
    V = verts * inst
    scene = zn.Scene(ctx, inst)
    scene.FillSynthetic(0x5A454E59 + 2, 1, 0.0)
    scene.UpdateWorld()
    # table=True: a ragged batch of the same mean size — every other boundary moved by one vertex — so that the kernel
    # really reads instance_first_vertex (an explicit table of EQUAL ranges is recognised and runs the arithmetic path)
    first = None
    if table:
        first = (np.arange(inst + 1, dtype=np.uint64) * verts).astype(np.uint32)
        first[1:-1:2] += 1
    mesh = zn.Mesh(ctx, V, inst, first)
    mesh.FillSynthetic(0x5A454E59 + 4)
    mesh.BindSceneModels(scene, 0)
    for _ in range(3):
        mesh.Transform()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(20):
        mesh.Transform()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    alg = V * 48 + inst * 64
    out[f"{inst}x{verts}" + ("/table" if table else "")] = {"vertices": V, "ms": ms, "G_vertices_per_s": V / (ms * 1e-3) / 1e9,
                              "frac_48B_per_vertex": V * 48 / (ms * 1e-3) / 1e9 / peak,
                              "frac_incl_model_matrices": alg / (ms * 1e-3) / 1e9 / peak}
    mesh.Close()
    scene.Close()
print(json.dumps(out, indent=1))
