"""One-GPU probe: device time of zn_scene_expand_visible for R copies of the bench scene's record
(what every rank runs per frame at N = R + 1), against the bytes it writes.  Not a bench line."""
import json
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
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
sc = zn.Scene(ctx, E, offs)
sc.FillSynthetic(0x5A454E59 + 3, 8, 0.0)
out = {}
for eye_z in (1500.0, 600.0, 0.0):
    sc.SetCamera(zn.Camera(zn.CameraDesc(eye=(0.0, 0.0, eye_z), center=(0.0, 0.0, eye_z - 1.0))))
    words = sc.DeviceBuffers().visibility_record_words
    R = 7
    recs = torch.zeros(R * words, dtype=torch.int32, device=dev)
    sc.BindVisibilityRecord(recs.data_ptr())
    sc.Update()
    ctx.Sync()
    n_vis = sc.VisibleCount()
    for r in range(1, R):
        recs[r * words:(r + 1) * words] = recs[:words]
    lists = torch.empty(R * E, dtype=torch.int32, device=dev)
    counts = torch.zeros(R, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    res = {"visible": n_vis, "fraction": n_vis / E}
    for mode, worklist in (("", False), ("worklist_", True)):
        sc.SetExpandMode(worklist)
        for r in (1, 3, 7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                sc.ExpandVisible(recs.data_ptr(), r, words, lists.data_ptr(), E, counts.data_ptr())
            a.record(stream)
            for _ in range(20):
                sc.ExpandVisible(recs.data_ptr(), r, words, lists.data_ptr(), E, counts.data_ptr())
            b.record(stream)
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / 20 * 1e3
            res[f"{mode}expand_{r}_us"] = us
            res[f"{mode}expand_{r}_write_GBs"] = r * n_vis * 4 / (us * 1e-6) / 1e9
    sc.SetExpandMode(False)
    # a plain device memset of the same bytes: what a pure write stream reaches here
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    view = lists[: 7 * n_vis]
    for _ in range(3):
        view.fill_(1)
    a.record(stream)
    for _ in range(20):
        view.fill_(1)
    b.record(stream)
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    res["fill_same_bytes_us"] = us
    res["fill_write_GBs"] = 7 * n_vis * 4 / (us * 1e-6) / 1e9
    sc.BindVisibilityRecord(None)
    out[f"eye_z_{int(eye_z)}"] = res
print(json.dumps(out, indent=1))
