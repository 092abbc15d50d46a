"""Smallest program that launches the round-2 kernels on the bench scene for ncu: world-only update, cull pass
(+ emit), and the multi-instance mesh tile kernel (4Mi instances x 8 vertices).  Not a bench."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zenyth_b200 as zn  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = zn.GpuContext(0, stream=stream.cuda_stream)
sizes = [7 * 8 ** l for l in range(8)]
E = sum(sizes)
sc = zn.Scene(ctx, E, np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32))
sc.FillSynthetic(0x5A454E59 + 3, 8, 0.0)
cam = zn.Camera(zn.CameraDesc(eye=(0.0, 0.0, 1500.0), center=(0.0, 0.0, 1499.0)))
for _ in range(3):
    sc.UpdateWorld()
    sc.Cull(cam)
ctx.Sync()
n = sc.VisibleCount()
inst, verts = 1 << 22, 8
flat = zn.Scene(ctx, inst)
flat.FillSynthetic(0x5A454E59 + 2, 1, 0.0)
flat.UpdateWorld()
mesh = zn.Mesh(ctx, inst * verts, inst)
mesh.FillSynthetic(0x5A454E59 + 4)
mesh.BindSceneModels(flat, 0)
for _ in range(3):
    mesh.Transform()
ctx.Sync()
print("ok", n)
