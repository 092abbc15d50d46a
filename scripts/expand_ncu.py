"""Smallest program that launches the expansion kernel on mixed records (camera inside the 16.7M-entity
scene): for `ncu -k regex:expand_visible`.  Not a bench."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zenyth_b200 as zn  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 7
EYE_Z = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = zn.GpuContext(0, stream=stream.cuda_stream)
sizes = [7 * 8 ** l for l in range(8)]
E = sum(sizes)
sc = zn.Scene(ctx, E, np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32))
sc.FillSynthetic(0x5A454E59 + 3, 8, 0.0)
sc.SetCamera(zn.Camera(zn.CameraDesc(eye=(0.0, 0.0, EYE_Z), center=(0.0, 0.0, EYE_Z - 1.0))))
sc.SetExpandMode(os.environ.get("ZN_EXPAND_MODE", "grid") == "worklist")
words = sc.DeviceBuffers().visibility_record_words
recs = torch.zeros(R * words, dtype=torch.int32, device=dev)
sc.BindVisibilityRecord(recs.data_ptr())
sc.Update()
ctx.Sync()
for r in range(1, R):
    recs[r * words:(r + 1) * words] = recs[:words]
lists = torch.empty(R * E, dtype=torch.int32, device=dev)
counts = torch.zeros(R, dtype=torch.int32, device=dev)
for _ in range(5):
    sc.ExpandVisible(recs.data_ptr(), R, words, lists.data_ptr(), E, counts.data_ptr())
torch.cuda.synchronize()
print("ok", counts.tolist())
