"""One-GPU probe of what a rank does per frame at N = 8, without peers: one Update() of the bench scene plus the
expansion of 7 copies of its record — on the same stream, or on a side stream (default / low priority) so that
the expansion can overlap the next frame.  Also: the cost of storing the record into 7 extra local targets from
inside the kernels (the instruction side of the peer stores; no NVLink involved).  Not a bench line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zenyth_b200 as zn  # noqa: E402

dev = torch.device("cuda", 0)
main = torch.cuda.Stream(dev, priority=-1)
torch.cuda.set_stream(main)
ctx = zn.GpuContext(0, stream=main.cuda_stream)
sizes = [7 * 8 ** l for l in range(8)]
E = sum(sizes)
sc = zn.Scene(ctx, E, np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32))
sc.FillSynthetic(0x5A454E59 + 3, 8, 0.0)
sc.SetCamera(zn.Camera(zn.CameraDesc(eye=(0.0, 0.0, 1500.0), center=(0.0, 0.0, 1499.0))))
words = sc.DeviceBuffers().visibility_record_words
R = 7
recs = torch.zeros((R + 1) * words, dtype=torch.int32, device=dev)
sc.BindVisibilityRecord(recs.data_ptr())
sc.Update()
ctx.Sync()
for r in range(1, R + 1):
    recs[r * words:(r + 1) * words] = recs[:words]
lists = torch.empty(R * E, dtype=torch.int32, device=dev)
counts = torch.zeros(R, dtype=torch.int32, device=dev)
torch.cuda.synchronize()


def timed(step, n=50):
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for _ in range(n):
        step()
    b.record(main)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def expand():
    sc.ExpandVisible(recs.data_ptr() + 4 * words, R, words, lists.data_ptr(), E, counts.data_ptr())


out = {"update_only_us": timed(sc.Update)}
out["expand_only_us"] = timed(expand)
# the signalled form with every flag already up: what the acquire (LDG.STRONG.SYS + CCTL.IVALL per CTA) costs
flags = torch.full((R * 32,), 1 << 20, dtype=torch.int32, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
torch.cuda.synchronize()


def expand_signalled():
    sc.ExpandVisibleSignalled(recs.data_ptr() + 4 * words, R, words, lists.data_ptr(), E, counts.data_ptr(), flags.data_ptr(), 32, 1,
                              status.data_ptr(), skip_index=-1, skip_count=0)


out["expand_signalled_flags_up_us"] = timed(expand_signalled)
assert status.item() == 0


def serial():
    sc.Update()
    expand()


out["update_then_expand_same_stream_us"] = timed(serial)
for name, prio in (("side_stream_default_priority", -1), ("side_stream_low_priority", 0)):
    side = torch.cuda.Stream(dev, priority=prio)
    done = [torch.cuda.Event(), torch.cuda.Event()]
    upd = torch.cuda.Event()
    k = [0]

    def overlapped():
        if k[0] >= 2:
            main.wait_event(done[k[0] % 2])
        sc.Update()
        upd.record(main)
        side.wait_event(upd)
        ctx.SetStream(side.cuda_stream)
        expand()
        ctx.SetStream(main.cuda_stream)
        done[k[0] % 2].record(side)
        k[0] += 1

    us = timed(overlapped)
    main.wait_stream(side)
    torch.cuda.synchronize()
    out[name + "_us"] = us
# the instruction side of the peer stores: 7 extra LOCAL record targets written from inside the kernels
extra = torch.zeros(7 * words, dtype=torch.int32, device=dev)
sc.BindPeerRecords([extra.data_ptr() + 4 * words * r for r in range(7)])
out["update_with_7_extra_local_record_targets_us"] = timed(sc.Update)
sc.BindPeerRecords([])
print(json.dumps(out, indent=1))
