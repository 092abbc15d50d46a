#!/usr/bin/env python
"""bench.py — the headline measurement: TRS->world + MVP + frustum cull + compaction, M entities/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload (BASELINE.json configs[2], the one the north-star target is quoted on): one synthetic
"H8-geo" scene per GPU — 7 roots, fan-out 8, 8 levels, 16,777,215 entities, level-ordered — seen by
the 'mid' camera of SURVEY.md §8(d).  A step is one zn_scene_update over that scene.  Inputs
(1.07 GB) and outputs (2.15 GB) are far larger than the 126 MB L2, so no flush is needed between
steps.  With N > 1 (torchrun, one rank per GPU) every rank owns one such scene (weak scaling) and
the step also all-gathers the visible counts and the compacted index lists over NCCL.

The JSON line carries, beside the driver contract: `roofline` (the dominant kernel — the last
hierarchy level — timed with CUDA events inside the library, against MEASURED_PEAKS.json),
`cpu_baseline` (the CPU oracle on the host cores, reported, not a target) and `e2e` (the same
metric through zn_scene_frame_host with pinned HOST buffers: per step H2D of position/rotation/
scale, D2H of the visible count + list).

`--impl reference` times the CPU statement of the path (oracle/, on the reference's own vec3/vec4
types when oracle/_ref is present) with all host threads.  The upstream engine has no
implementation of this path to run, so the oracle IS the reference arm.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASE_SEED = 0x5A454E59
METRIC = "entities_per_second_trs_world_mvp_cull"
UNIT = "M entities/s"

WORKLOADS = {
    # name: (roots, levels, fanout, description)
    "h8geo16m": (7, 8, 8, "H8-geo: 7 roots, fan-out 8, 8 levels, 16,777,215 entities (BASELINE configs[2])"),
    "flat1m": (1048576, 1, 1, "flat: 1,048,576 root entities (BASELINE configs[1])"),
    "h8geo32m": (14, 8, 8, "H8-geo: 14 roots, fan-out 8, 8 levels, 33,554,430 entities (one scene of BASELINE configs[4])"),
    "h8chain16m": (2097152, 8, 1, "H8-chain: 8 levels x 2,097,152, fan-out 1 (SURVEY §8d stress variant)"),
}
CAMERA_EYE = (0.0, 0.0, 1500.0)


_JSON_OUT = None


def capture_stdout() -> None:
    """The driver wants ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    line on fd 1), so fd 1 is pointed at stderr for the run and the JSON goes to the real stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit_line(text: str, flush: bool = True) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def level_sizes(roots, levels, fanout):
    return [roots * fanout ** l for l in range(levels)]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id: str):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", gpu_id, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


class NvmlSampler:
    """SM clock / throttle reasons polled through NVML back to back (0.5 ms pause) from a thread DURING
    the timed region (the timed region is ~10 ms: nvidia-smi's 100 ms loop sees it once or twice)."""

    def __init__(self, gpu_uuid: str):
        import threading

        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByUUID(gpu_uuid.encode() if isinstance(gpu_uuid, str) else gpu_uuid)
        self.sm, self.power, self.reasons = [], [], 0
        self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.0005)

    def stop(self) -> dict:
        self._stop.set()
        self._t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        active = sorted(k for k, bit in names.items() if self.reasons & int(bit))
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm, "power_w_max": float(max(self.power)) if self.power else None,
                "samples": len(self.sm), "reasons": active, "source": "NVML polled back to back (0.5 ms pause) during the timed region"}


def make_sampler(gpu_uuid: str):
    try:
        return NvmlSampler(gpu_uuid)
    except Exception:
        return ClockSampler(gpu_uuid)


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle, all host threads, preallocated buffers (first-touch done in warm-up)
class CpuScene:
    def __init__(self, workload: str, seed: int, reference_types: bool = False):
        from oracle.binding import Oracle, _ptr, camera
        self._ptr = _ptr
        self.oracle = Oracle(reference_types=reference_types)
        self.kind_note = ("oracle built on the reference's own zenyth::math vec3/vec4/mat4 (oracle/_ref; out-of-line SSE calls)"
                          if reference_types else
                          "oracle with the restated, inlined vec3/vec4/mat4 (the faster of the two oracle builds; "
                          "bit-identical results to the oracle/_ref build on the reference's own types)")
        roots, levels, fanout, _ = WORKLOADS[workload]
        self.threads = self.oracle.hardware_threads()
        self.arrays = self.oracle.synth_scene(seed, level_sizes(roots, levels, fanout), fanout, 0.0, self.threads)
        self.n = int(self.arrays["position"].shape[0])
        self.view, self.proj = camera(self.oracle, eye=CAMERA_EYE)
        self.world = np.zeros((self.n, 16), np.float32)
        self.mvp = np.zeros((self.n, 16), np.float32)
        self.flags = np.zeros(self.n, np.uint8)
        self.visible = np.zeros(self.n, np.uint32)
        self.count = C.c_uint32(0)

    def frame(self):
        a, o, p = self.arrays, self.oracle, self._ptr
        lo = a["level_offsets"]
        o.lib.zo_scene_update(C.c_uint32(self.n), p(lo, np.uint32), C.c_uint32(lo.size - 1),
                              p(a["position"], np.float32), p(a["rotation"], np.float32), p(a["scale"], np.float32),
                              p(a["aabb_center"], np.float32), p(a["aabb_half"], np.float32), p(a["parent"], np.uint32),
                              o._v(self.view, 16), o._v(self.proj, 16),
                              p(self.world, np.float32), p(self.mvp, np.float32), p(self.flags, np.uint8),
                              p(self.visible, np.uint32), C.byref(self.count), C.c_uint(self.threads))
        return self.count.value


def cpu_baseline(workload: str, budget_s: float = 12.0) -> dict:
    sc = CpuScene(workload, BASE_SEED + 3)
    sc.frame()   # first touch
    frames, t0 = 0, time.perf_counter()
    while True:
        sc.frame()
        frames += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or frames >= 50:
            break
    out = {"value": sc.n * frames / dt / 1e6, "unit": UNIT, "cores": sc.threads, "kind": "port",
           "sample": f"{frames} frames of the full {sc.n:,}-entity {workload} scene, {sc.threads} host threads "
                     f"splitting each level by entity range; {sc.kind_note}",
           "ms_per_frame": dt / frames * 1e3}
    all_threads, sc.threads = sc.threads, 1          # SURVEY.md §8d also asks for the scalar single-thread figure
    t0 = time.perf_counter()
    sc.frame()
    out["one_thread"] = {"value": sc.n / (time.perf_counter() - t0) / 1e6, "unit": UNIT, "cores": 1, "sample": "1 frame of the same scene"}
    sc.threads = all_threads
    del sc
    from oracle.binding import LIB_REFERENCE
    if os.path.exists(LIB_REFERENCE):
        rs = CpuScene(workload, BASE_SEED + 3, reference_types=True)
        rs.frame()
        t0 = time.perf_counter()
        rs.frame()
        out["on_reference_types"] = {"value": rs.n / (time.perf_counter() - t0) / 1e6, "unit": UNIT,
                                     "note": "same oracle on the reference's own vec3/vec4/mat4 (oracle/_ref), 1 frame"}
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = CpuScene(args.workload, BASE_SEED + 3)
    for _ in range(max(args.warmup, 1)):
        sc.frame()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        count = sc.frame()
    dt = time.perf_counter() - t0
    value = sc.n * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][3], "entities": sc.n, "camera_eye": CAMERA_EYE,
                   "visible_fraction": count / sc.n, "where": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": sc.threads, "kind": "port",
                         "sample": f"{args.steps} frames of the full {sc.n:,}-entity scene per run; {sc.kind_note}. "
                                   "The upstream engine has no implementation of this path (SURVEY.md §0), so the oracle is the CPU arm."},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
class DevArray:
    """Zero-copy view of a library-owned device buffer for torch (NCCL plumbing only)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import zenyth_b200 as zn

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world_size > 1
    if distributed:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)

    roots, levels, fanout, desc = WORKLOADS[args.workload]
    sizes = level_sizes(roots, levels, fanout)
    E = sum(sizes)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)

    # One explicit stream for everything: torch's current stream (events, NCCL) == the library's.
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = zn.GpuContext(local_rank, stream=stream.cuda_stream)
    scene = zn.Scene(ctx, E, offsets if levels > 1 else None)
    scene.FillSynthetic(BASE_SEED + 3 + rank, fanout, 0.0)
    cam = zn.Camera(zn.CameraDesc(eye=CAMERA_EYE, center=(CAMERA_EYE[0], CAMERA_EYE[1], CAMERA_EYE[2] - 1.0)))
    scene.SetCamera(cam)
    scene.SetIndexBase(rank * E if distributed and rank * E < 2 ** 32 else 0)
    bufs = scene.DeviceBuffers()
    launches_per_update = scene.LaunchesPerUpdate()

    frame_no = [0]
    mode = args.exchange if distributed else "none"
    fallback_note = ""
    index_bases = np.asarray([(r * E) % 2 ** 32 for r in range(world_size)], np.uint32)
    if mode.startswith("peer"):
        # The visibility records (1 bit per entity), but no collective moves them: the level and emit
        # kernels store every tile's mask and the record header straight into all peers' buffers over
        # NVLink (CUDA IPC mapped peer memory); completion flags travel the same way and the
        # expansion kernel waits on them itself ("peer-nccl": a 4-byte all-reduce per frame is the
        # cross-rank barrier instead; "peer-waitkernel": signal + wait as a launch of their own).
        from zenyth_b200.peer_exchange import PeerRecordExchange
        try:
            exchange = PeerRecordExchange(ctx, bufs.visibility_record_words, E, world_size, rank, dev, torch.int32, slots=4,
                                          signal="nccl" if mode == "peer-nccl" else "flags", fused_wait=mode != "peer-waitkernel")
        except RuntimeError as e:    # raised on every rank together: peer mapping is not possible on this box
            fallback_note = f"--exchange {mode} unavailable ({e}); ran --exchange records"
            mode = "records"
    use_lists, use_records, use_peer = mode == "lists", mode == "records", mode.startswith("peer")
    if use_lists:
        # The exchange of the path as SURVEY §8e words it: all-gather the visible counts, then the
        # compacted index lists.  Double-buffered: the scene writes frame k's list straight into
        # slot k%2 of the exchange, whose NCCL all-gathers run on the collective's stream while
        # frame k+1 is computed; a slot is collected (checked, exact) before it is reused.
        from zenyth_b200.distributed import VisibleExchange
        exchange = VisibleExchange(E, world_size, dev, torch.int32, slots=2)
    if use_records:
        # Same result, 1 bit per entity on the wire: all-gather the fixed-size visibility records,
        # every rank expands all of them into the ordered lists (zenyth_b200.distributed.RecordExchange).
        from zenyth_b200.distributed import RecordExchange
        exchange = RecordExchange(bufs.visibility_record_words, E, world_size, dev, torch.int32, slots=2)

    def step():
        slot = frame_no[0] % 2
        if use_lists:
            exchange.acquire(slot)
            scene.BindVisibleOutput(exchange.local_ptr(slot), E, exchange.count_ptr(slot))
            scene.Update()
            exchange.launch(slot)
        elif use_records:
            scene.BindVisibilityRecord(exchange.record_ptr(slot))
            scene.Update()
            exchange.launch(slot)                      # all-gather of frame k: overlaps what follows
            if exchange.ready(1 - slot):               # frame k-1 has arrived: derive every rank's list
                exchange.expand(scene, 1 - slot, index_bases)
        elif use_peer:
            k = frame_no[0]
            exchange.bind(scene, k)
            scene.Update()                             # kernels write this rank's record into every peer, then signal
            exchange.launch(k)                         # (peer-nccl only: barrier(k), asynchronous)
            exchange.expand(scene, k - 1, index_bases)   # waits on the device for the peers' frame k-1 signals
        else:
            scene.Update()
        frame_no[0] += 1

    def finish():
        """All frames issued so far are complete, including their exchanges."""
        if use_lists:
            exchange.drain()
        if use_records:
            for slot in ((frame_no[0] - 2) % 2, (frame_no[0] - 1) % 2):
                if exchange.ready(slot):
                    exchange.expand(scene, slot, index_bases)
        if use_peer:
            for k in (frame_no[0] - 2, frame_no[0] - 1):
                exchange.expand(scene, k, index_bases)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    finish()
    barrier()
    if use_lists:
        scene.BindVisibleOutput(None)
        scene.Update()
    visible = scene.VisibleCount()
    if use_records or use_peer:   # every rank derived the same lists: check them against each rank's own count
        counts, _ = exchange.result()
        own = torch.tensor([visible], dtype=torch.int32, device=dev)
        every = torch.zeros(world_size, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(every, own)
        assert counts == every.tolist(), (counts, every.tolist())
        # ... and every derived list against its owner's actual list (checksums travel, not lists)
        _, lists = exchange.result()
        mine = torch.from_numpy(scene.Visible().astype(np.int64)).to(dev)

        def checksum(t):
            t = t.to(torch.int64) & 0xFFFFFFFF
            return torch.stack([t.sum(), (t * (torch.arange(t.numel(), device=dev) % 1021 + 1)).sum()])

        sums = torch.zeros(world_size * 2, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, checksum(mine))
        for r in range(world_size):
            assert torch.equal(checksum(lists[r, : counts[r]]), sums[2 * r: 2 * r + 2]), f"rank {rank}: list of rank {r} differs from its owner's"

    gpu_id = str(torch.cuda.get_device_properties(dev).uuid)
    if not gpu_id.startswith("GPU-"):
        gpu_id = "GPU-" + gpu_id
    sampler = make_sampler(gpu_id)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(stream)
    for _ in range(args.steps):
        step()
    finish()                  # the stream now also waits for the last exchanges
    stop.record(stream)
    barrier()
    clocks = sampler.stop()
    if use_lists:
        scene.BindVisibleOutput(None)
    if use_records:
        scene.BindVisibilityRecord(None)
    if use_peer:
        exchange.close(scene)
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())

    # ---- dominant kernel: last (largest) level, timed inside the library with CUDA events --------
    scene.SetProfiling(True)
    level_ms = []
    tails = []
    for _ in range(args.steps):
        scene.Update()
        lm, tail = scene.LevelTimesMs()
        level_ms.append(lm)
        tails.append(tail)
    scene.SetProfiling(False)
    level_ms = np.asarray(level_ms, np.float64)
    dom = int(np.argmax(level_ms.mean(axis=0)))
    dom_ms = float(level_ms[:, dom].mean())
    dom_entities = int(sizes[dom])
    dom_bytes = dom_entities * (188 if dom == 0 else 192)
    peak, peak_src = measured_peaks()
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    frame_bytes = scene.AlgorithmicBytes(visible)
    frame_ms_prof = float(level_ms.sum(axis=1).mean() + np.mean(tails))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")

    # ---- e2e: the C-ABI call a CPU engine swaps in, HOST buffers, copies inside the timed region --
    e2e = None
    if not args.no_e2e:
        inputs = scene.Inputs()
        h_pos, h_rot, h_scl = (ctx.HostAlloc((E, 3), np.float32) for _ in range(3))
        h_pos[:] = inputs["position"]; h_rot[:] = inputs["rotation"]; h_scl[:] = inputs["scale"]
        del inputs
        h_vis = [ctx.HostAlloc((E,), np.uint32) for _ in range(2)]
        e2e_steps = args.steps
        # (1) one frame at a time: upload, update, download, return (zn_scene_frame_host)
        for _ in range(2):
            n_vis = scene.FrameHost(h_pos, h_rot, h_scl, visible=h_vis[0])
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            n_vis = scene.FrameHost(h_pos, h_rot, h_scl, visible=h_vis[0])
        torch.cuda.synchronize(dev)
        dt_sync = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        # (2) the same frames, two in flight (zn_scene_frame_host_submit / _wait): every step still
        # uploads its 36 B/entity and downloads its count + list; frame k+1's upload overlaps frame
        # k's update and download.  Timed over e2e_steps completed frames in steady state.
        scene.SubmitFrameHost(h_pos, h_rot, h_scl)
        scene.SubmitFrameHost(h_pos, h_rot, h_scl)
        for k in range(2):
            n_pipe = scene.WaitFrameHost(h_vis[k % 2])
            scene.SubmitFrameHost(h_pos, h_rot, h_scl)
        barrier_host = dist.barrier if distributed else (lambda: None)
        barrier_host()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            n_pipe = scene.WaitFrameHost(h_vis[k % 2])
            scene.SubmitFrameHost(h_pos, h_rot, h_scl)
        dt_pipe = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        for k in range(2):
            scene.WaitFrameHost(h_vis[k % 2])
        assert n_pipe == n_vis, (n_pipe, n_vis)
        if distributed:
            dist.all_reduce(dt_sync, op=dist.ReduceOp.MAX)
            dist.all_reduce(dt_pipe, op=dist.ReduceOp.MAX)
        e2e = {"value": world_size * E * e2e_steps / float(dt_pipe.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(E) * 36, "d2h_bytes_per_step": int(n_vis) * 4 + 4,
               "ms_per_step": float(dt_pipe.item()) / e2e_steps * 1e3,
               "api": "zn_scene_frame_host_submit/_wait, two frames in flight: every step uploads pinned host "
                      "position/rotation/scale and downloads the visible count + index list; world/MVP stay "
                      "device-resident for the renderer",
               "one_frame_at_a_time": {"value": world_size * E * e2e_steps / float(dt_sync.item()) / 1e6, "unit": UNIT,
                                       "ms_per_step": float(dt_sync.item()) / e2e_steps * 1e3, "api": "zn_scene_frame_host"},
               "timer": "host wall clock around the calls, steady state"}

    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.workload)

    if rank == 0:
        value = world_size * E * args.steps / (total_ms * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "entities_per_gpu": E, "levels": levels, "camera_eye": CAMERA_EYE,
                       "visible_fraction": visible / E,
                       "l2": "inputs 1.07 GB + outputs 2.15 GB per step >> 126 MB L2, no flush needed" if E > 4_000_000
                             else "footprint ~%d MB vs 126 MB L2" % (E * 192 // 1_000_000),
                       "exchange": mode,
                       "collective": (
                           ("NCCL all_gather of visible counts + index lists every step, double-buffered: the exchange of "
                            "frame k overlaps the update of frame k+1; the timed region ends after the last exchange; "
                            f"re-gathers after a size misprediction: {exchange.regathers}") if use_lists else
                           ("NCCL all_gather of the fixed-size visibility records (per-chunk counts + 1 mask bit per entity) "
                            "every step, double-buffered; every rank then expands all records into the ordered index lists + "
                            "counts (identical to an all-gather of the lists, 27x fewer NVLink bytes); the timed region "
                            "ends after the last expansion; --exchange lists runs the literal list all-gather") if use_records else
                           ("fused into the compute kernels over NVLink peer memory, no collective: the level and emit kernels store "
                            "the visibility record tile by tile into every peer's buffer (CUDA IPC mapped pointers); "
                            + ("the kernel that follows the update on the stream raises a sequence flag at every peer (one fence.sys + relaxed sys-scope stores) and the "
                               "expansion kernel waits on the owners' flags itself — no NCCL call in the timed region; "
                               if mode != "peer-nccl" else "a 4-byte NCCL all-reduce per frame is the cross-rank barrier; ")
                            + "every rank expands its peers' records into the ordered index lists + counts (its own list comes "
                              "straight from its emit kernel)") if use_peer
                           else "none (1 GPU)") + ((" [" + fallback_note + "]") if fallback_note else "")},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "frac can exceed 1: the peak is torch's copy_ (1 read : 1 write); this kernel moves 1 read : 1.73 write "
                                 "in 16 KiB TMA bursts and reaches 83% of the 8.18 TB/s ncu reports as the device's DRAM peak "
                                 "(profiles/r01_s2_level7_ncu_summary.csv), against 79% for the copy",
                         "kernel": f"scene_level_kernel<{'true' if dom else 'false'}> level {dom} ({dom_entities:,} entities)",
                         "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": dom_ms,
                         "frame": {"algorithmic_bytes": frame_bytes, "ms": total_ms / args.steps / (1 if not distributed else 1),
                                   "achieved": frame_bytes / (total_ms / args.steps * 1e-3) / 1e9,
                                   "frac": frame_bytes / (total_ms / args.steps * 1e-3) / 1e9 / peak,
                                   "level_ms": [float(x) for x in level_ms.mean(axis=0)], "scan_emit_ms": float(np.mean(tails)),
                                   "profiled_frame_ms": frame_ms_prof}},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": args.steps * (launches_per_update + (1 if (use_records or use_peer) else 0)),
        }
        emit_line(json.dumps(line), flush=True)
    scene.Close()
    ctx.Close()
    if distributed:
        dist.destroy_process_group()


def run_mesh(args) -> None:
    """Secondary measurement (BASELINE configs[3]): 65,536 instances x 1,024 vertices = 64Mi vertices,
    position + normal transform by the world matrices of a 65,536-entity flat scene.  M vertices/s."""
    import torch

    import zenyth_b200 as zn
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = zn.GpuContext(0, stream=stream.cuda_stream)
    inst, verts = 65536, 1024
    V = inst * verts
    scene = zn.Scene(ctx, inst)
    scene.FillSynthetic(BASE_SEED + 2, 1, 0.0)
    scene.SetCamera(zn.Camera(zn.CameraDesc(eye=CAMERA_EYE, center=(0.0, 0.0, CAMERA_EYE[2] - 1.0))))
    scene.Update()
    mesh = zn.Mesh(ctx, V, inst)
    mesh.FillSynthetic(BASE_SEED + 4)
    mesh.BindSceneModels(scene, 0)
    for _ in range(max(args.warmup, 3)):
        mesh.Transform()
    torch.cuda.synchronize(dev)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for _ in range(args.steps):
        mesh.Transform()
    stop.record(stream)
    torch.cuda.synchronize(dev)
    ms = start.elapsed_time(stop) / args.steps
    peak, peak_src = measured_peaks()
    achieved = V * 48 / (ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu:
        from oracle.binding import Oracle
        o = Oracle()
        th = o.hardware_threads()
        n_i = 8192
        pos, nrm = o.synth_mesh(BASE_SEED + 4, n_i * verts, th)
        models = scene.World(0, n_i)
        first = (np.arange(n_i + 1, dtype=np.uint32) * verts).astype(np.uint32)
        o.mesh_transform(first, models, pos, nrm, th)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 5.0:
            o.mesh_transform(first, models, pos, nrm, th)
            reps += 1
        cpu = {"value": n_i * verts * reps / (time.perf_counter() - t0) / 1e6, "unit": "M vertices/s", "cores": th, "kind": "port",
               "sample": f"{reps} passes over {n_i} instances x {verts} vertices (1/8 of the batch), {th} host threads"}
    emit_line(json.dumps({
        "metric": "vertices_per_second_mesh_transform", "value": V / (ms * 1e-3) / 1e6, "unit": "M vertices/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mesh batch: 65,536 instances x 1,024 vertices = 67,108,864 vertices (BASELINE configs[3])",
                   "l2": "3.2 GB touched per step >> 126 MB L2, no flush needed"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "mesh_tile_kernel", "algorithmic_bytes_per_launch": V * 48, "launch_ms": ms},
        "cpu_baseline": cpu, "gpu_launches": args.steps}), flush=True)
    mesh.Close()
    scene.Close()
    ctx.Close()


def run_scenes8(args) -> None:
    """BASELINE configs[4]: 8 independent 33,554,430-entity H8-geo scenes sharded over N = 1/2/4/8
    GPUs (8/N scenes per GPU, updated back to back), then the exchange: every rank ends up with all
    8 compacted visible lists and counts.  Strong scaling: the total work is fixed."""
    import torch
    import torch.distributed as dist

    import zenyth_b200 as zn
    from zenyth_b200.distributed import RecordExchange, scene_index_bases, shard_scenes

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world_size > 1
    if distributed:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)
    total_scenes = 8
    assert total_scenes % world_size == 0
    owned = shard_scenes(total_scenes, world_size, rank)
    sizes = level_sizes(14, 8, 8)
    E = sum(sizes)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    bases = scene_index_bases([E] * total_scenes)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = zn.GpuContext(local_rank, stream=stream.cuda_stream)
    cam = zn.Camera(zn.CameraDesc(eye=CAMERA_EYE, center=(CAMERA_EYE[0], CAMERA_EYE[1], CAMERA_EYE[2] - 1.0)))
    scenes = []
    for sid in owned:
        sc = zn.Scene(ctx, E, offsets)
        sc.FillSynthetic(BASE_SEED + 5 + sid, 8, 0.0)
        sc.SetCamera(cam)
        sc.SetIndexBase(bases[sid])
        scenes.append(sc)
    words = scenes[0].DeviceBuffers().visibility_record_words
    launches_per_update = scenes[0].LaunchesPerUpdate()
    frame_no = [0]
    mode, fallback_note = (args.exchange if distributed else "none"), ""
    if mode == "lists":
        raise SystemExit("bench.py: --workload scenes8x32m exchanges visibility records (--exchange peer | records | peer-nccl | peer-waitkernel)")
    index_bases = np.asarray(bases, np.uint32)
    if mode.startswith("peer"):
        from zenyth_b200.peer_exchange import PeerRecordExchange
        try:
            exchange = PeerRecordExchange(ctx, words, E, world_size, rank, dev, torch.int32, slots=4, records_per_rank=len(owned),
                                          signal="nccl" if mode == "peer-nccl" else "flags", fused_wait=mode != "peer-waitkernel")
        except RuntimeError as e:    # raised on every rank together
            fallback_note = f" [--exchange {mode} unavailable ({e}); ran --exchange records]"
            mode = "records"
    use_peer = mode.startswith("peer")
    if mode == "records":
        exchange = RecordExchange(words, E, world_size, dev, torch.int32, slots=2, records_per_rank=len(owned))

    def step():
        k = frame_no[0]
        if use_peer:
            exchange.bind(scenes, k)
            for sc in scenes:
                sc.Update()                              # kernels store the records into every peer as they run
            exchange.launch(k)
            exchange.expand(scenes[0], k - 1, index_bases)
        else:
            slot = k % 2
            for j, sc in enumerate(scenes):
                if distributed:
                    sc.BindVisibilityRecord(exchange.record_ptr(slot, j))
                sc.Update()
            if distributed:
                exchange.launch(slot)
                if exchange.ready(1 - slot):
                    exchange.expand(scenes[0], 1 - slot, index_bases)
        frame_no[0] += 1

    def finish():
        if use_peer:
            for k in (frame_no[0] - 2, frame_no[0] - 1):
                exchange.expand(scenes[0], k, index_bases)
        elif distributed:
            for slot in ((frame_no[0] - 2) % 2, (frame_no[0] - 1) % 2):
                if exchange.ready(slot):
                    exchange.expand(scenes[0], slot, index_bases)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    finish()
    barrier()
    local_counts = [sc.VisibleCount() for sc in scenes]
    if distributed:
        # every rank derived all 8 lists: check counts and list checksums against the owners' own lists
        counts, lists = exchange.result()

        def checksum(t):
            t = t.to(torch.int64) & 0xFFFFFFFF
            return torch.stack([t.sum(), (t * (torch.arange(t.numel(), device=dev) % 1021 + 1)).sum()])

        own = torch.stack([torch.cat([torch.tensor([n], dtype=torch.int64, device=dev),
                                      checksum(torch.from_numpy(sc.Visible().astype(np.int64)).to(dev))])
                           for sc, n in zip(scenes, local_counts)]).flatten()
        every = torch.zeros(total_scenes * 3, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, own)
        every = every.view(total_scenes, 3)
        assert counts == every[:, 0].tolist(), (counts, every[:, 0].tolist())
        for sid in range(total_scenes):
            assert torch.equal(checksum(lists[sid, : counts[sid]]), every[sid, 1:]), f"rank {rank}: list of scene {sid} differs from its owner's"
    else:
        counts = local_counts
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(stream)
    for _ in range(args.steps):
        step()
    finish()
    stop.record(stream)
    barrier()
    if use_peer:
        exchange.close(scenes)
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    if rank == 0:
        peak, peak_src = measured_peaks()
        alg = total_scenes * (14 * 188 + (E - 14) * 192) + 4 * sum(counts)
        emit_line(json.dumps({
            "metric": METRIC, "value": total_scenes * E * args.steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "8 independent H8-geo scenes of 33,554,430 entities sharded over the GPUs (BASELINE configs[4])",
                       "scenes_per_gpu": len(owned), "visible_counts": counts,
                       "exchange": mode,
                       "collective": (("visibility records stored by the level/emit kernels straight into every peer's memory over NVLink "
                                       "(CUDA IPC), completion flags in peer memory, local expansion of the peers' lists on every rank; "
                                       "no collective in the timed region") if mode in ("peer", "peer-waitkernel") else
                                      "peer stores of the records + a 4-byte NCCL all-reduce per frame as the barrier" if mode == "peer-nccl" else
                                      "NCCL all_gather of visibility records + local expansion of all 8 lists on every rank" if distributed
                                      else "none (1 GPU: every scene's own emit already leaves its list on the device)") + fallback_note},
            "roofline": {"bound": "hbm", "achieved": alg / (total_ms / args.steps * 1e-3) / 1e9 / world_size, "peak": peak, "unit": "GB/s",
                         "frac": alg / (total_ms / args.steps * 1e-3) / 1e9 / world_size / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "whole step, per GPU (algorithmic bytes of all 8 frames / step time / N)"},
            "gpu_launches": args.steps * (len(owned) * launches_per_update + (1 if distributed else 0))}), flush=True)
    for sc in scenes:
        sc.Close()
    ctx.Close()
    if distributed:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["mesh64m", "scenes8x32m"], default="h8geo16m")
    ap.add_argument("--exchange", choices=["peer", "records", "lists", "peer-nccl", "peer-waitkernel"], default="peer",
                    help="N>1: the kernels store the visibility records straight into peer GPUs' memory over NVLink and signal "
                         "completion with flags in peer memory (peer, default; falls back to records where peer mapping is refused); "
                         "NCCL all-gather of the records (records) or of the index lists (lists); peer stores with a 4-byte "
                         "all-reduce as the barrier (peer-nccl) or with signal + wait as a kernel of their own (peer-waitkernel)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg")
    args = ap.parse_args()
    capture_stdout()
    if args.workload == "mesh64m":
        run_mesh(args)
    elif args.workload == "scenes8x32m":
        run_scenes8(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
