"""A/B of zn_scene_set_level_fusion (scene_top_kernel) on the GPU: back-to-back frames (device time per frame) and
host-synchronised frames (latency per frame) for the bench scene, the config-1 stand-in and a mid-sized tree.
Prints one JSON object.  Usage: python scripts/level_fusion_probe.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import zenyth_b200 as zn


def main():
    out = {}
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    ctx = zn.GpuContext(0, stream=stream.cuda_stream)
    shapes = {"h8geo16m (7 roots, 8 levels)": [7 * 8 ** l for l in range(8)],
              "config-1 stand-in (4,681 entities, 5 levels)": [8 ** l for l in range(5)],
              "h8geo 6 levels (262k entities)": [7 * 8 ** l for l in range(6)]}
    for name, sizes in shapes.items():
        E = sum(sizes)
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
        sc = zn.Scene(ctx, E, offsets)
        sc.FillSynthetic(0x5A454E59 + 3, 8, 0.0)
        sc.SetCamera(zn.Camera(zn.CameraDesc(eye=(0.0, 0.0, 1500.0), center=(0.0, 0.0, 1499.0))))
        rec = {}
        for world_only in (False, True):
            for fusion in (True, False, True, False):
                sc.SetLevelFusion(fusion)
                step = sc.UpdateWorld if world_only else sc.Update
                for _ in range(20):
                    step()
                ctx.Sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 200
                e0.record(stream)
                for _ in range(n):
                    step()
                e1.record(stream)
                ctx.Sync()
                back_to_back = e0.elapsed_time(e1) / n * 1e3
                t0 = time.perf_counter()
                for _ in range(n):
                    step()
                    ctx.Sync()
                synced = (time.perf_counter() - t0) / n * 1e6
                key = ("update_world" if world_only else "update") + ("/fused" if fusion else "/per-level")
                rec.setdefault(key, []).append({"us_per_frame_back_to_back": round(back_to_back, 2), "us_per_frame_host_synchronised": round(synced, 2)})
        out[name] = rec
        sc.Close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
