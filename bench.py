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


FLAT16M = 16_777_216


def l2_note(E: int) -> str:
    return ("inputs 1.07 GB + outputs 2.15 GB per step >> 126 MB L2, no flush needed" if E > 4_000_000
            else "footprint ~%d MB vs 126 MB L2" % (E * 192 // 1_000_000))


def workload_config(workload: str, E: int, levels: int, visible: int) -> dict:
    """The `config` object — the SAME keys and values in both arms (ours and --impl reference), so the driver's
    same-config check compares like with like; everything arm-specific lives in other top-level keys."""
    return {"workload": WORKLOADS[workload][3], "entities_per_gpu": int(E), "levels": int(levels), "camera_eye": list(CAMERA_EYE),
            "visible_fraction": visible / E, "l2": l2_note(E)}


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
        "config": workload_config(args.workload, sc.n, WORKLOADS[args.workload][1], count),
        "where": "host CPU",
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


class Rig:
    """One process = one GPU: device, the ONE stream everything runs on (torch's current stream == the
    library's), the context, and torch.distributed for launch / timing plumbing only."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import zenyth_b200 as zn
        self.torch, self.dist, self.zn = torch, dist, zn
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.distributed = self.world_size > 1
        self.numa = bind_to_gpu_numa_node(self.local_rank)      # before any pinned allocation: first touch lands on the GPU's node
        if self.distributed:
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version there)
            dist.init_process_group("nccl", device_id=self.dev)
        # high priority: the exchange's side stream (lowest priority) then yields SM slots to the frame's kernels
        self.stream = torch.cuda.Stream(self.dev, priority=-1)
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.ctx = zn.GpuContext(self.local_rank, stream=self.stream.cuda_stream)
        gpu_id = str(torch.cuda.get_device_properties(self.dev).uuid)
        self.gpu_id = gpu_id if gpu_id.startswith("GPU-") else "GPU-" + gpu_id

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps: int, finish=None) -> float:
        """ms for `steps` calls of fn (+ finish), CUDA events on the stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        start.record(self.stream)
        for _ in range(steps):
            fn()
        if finish:
            finish()
        stop.record(self.stream)
        self.barrier()
        ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def close(self):
        self.ctx.Close()
        if self.distributed:
            self.dist.destroy_process_group()


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so the pinned host buffers of the e2e leg
    (first-touched below) and the thread that enqueues the copies are node-local.  Best effort; reports what it did."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                                              int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None, "note": "the platform reports no NUMA node for this GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"node": node, "note": "none of the node's CPUs is available to this process"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed), "note": "process pinned to the GPU's NUMA node before allocating pinned host memory"}
    except Exception as e:      # noqa: BLE001 - best effort
        return {"node": None, "note": f"not bound ({type(e).__name__}: {e})"}


def h2d_ceiling(rig: Rig, nbytes: int) -> dict:
    """What this box's host->device path delivers for the e2e upload: one pinned buffer of the upload's size copied
    with cudaMemcpyAsync on every rank AT THE SAME TIME (best of 5, device-timed, max over ranks)."""
    torch = rig.torch
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.zero_()
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=rig.dev)
    best = None
    for _ in range(6):
        ms = rig.timed(lambda: devbuf.copy_(host, non_blocking=True), 1)
        best = ms if best is None else min(best, ms)
    del host, devbuf
    return {"per_gpu_GBs": nbytes / (best * 1e-3) / 1e9, "aggregate_GBs": rig.world_size * nbytes / (best * 1e-3) / 1e9,
            "bytes": nbytes, "how": "torch copy_ of one pinned buffer of the upload's size, all ranks at once, best of 6"}


def profiled_levels(scene, steps: int):
    """(level_ms[steps, L], tail_ms[steps]) measured with CUDA events inside the library."""
    scene.SetProfiling(True)
    level_ms, tails = [], []
    for _ in range(steps):
        scene.Update()
        lm, tail = scene.LevelTimesMs()
        level_ms.append(lm)
        tails.append(tail)
    scene.SetProfiling(False)
    return np.asarray(level_ms, np.float64), np.asarray(tails, np.float64)


def expand_probe(rig: Rig, scene, E: int, visible: int, records: int = 7, reps: int = 20) -> dict:
    """What a rank does per frame at N = records + 1: expand `records` copies of this frame's visibility record into
    index lists — against a plain fill of the same bytes on the same GPU (the pure-write ceiling)."""
    torch = rig.torch
    words = scene.DeviceBuffers().visibility_record_words
    recs = torch.zeros(records * words, dtype=torch.int32, device=rig.dev)
    scene.BindVisibilityRecord(recs.data_ptr())
    scene.Update()
    rig.ctx.Sync()
    for r in range(1, records):
        recs[r * words:(r + 1) * words] = recs[:words]
    lists = torch.empty(records * E, dtype=torch.int32, device=rig.dev)
    counts = torch.zeros(records, dtype=torch.int32, device=rig.dev)
    call = lambda: scene.ExpandVisible(recs.data_ptr(), records, words, lists.data_ptr(), E, counts.data_ptr())   # noqa: E731
    for _ in range(3):
        call()
    us = rig.timed(call, reps) / reps * 1e3
    assert counts.tolist() == [visible] * records
    view = lists[: records * visible]
    for _ in range(3):
        view.fill_(1)
    fill_us = rig.timed(lambda: view.fill_(1), reps) / reps * 1e3
    scene.SetExpandMode(True)                   # the two-launch form: classification + work list (ZN_EXPAND_WORKLIST)
    for _ in range(3):
        call()
    us_wl = rig.timed(call, reps) / reps * 1e3
    assert counts.tolist() == [visible] * records
    scene.SetExpandMode(False)
    scene.BindVisibilityRecord(None)
    del recs, lists
    return {"records": records, "expand_us": us, "list_write_GBs": records * visible * 4 / (us * 1e-6) / 1e9,
            "fill_same_bytes_us": fill_us, "frac_of_fill": fill_us / us,
            "worklist_expand_us": us_wl, "worklist_frac_of_fill": fill_us / us_wl}


def secondary_lines(rig: Rig, scene, E: int, roots: int, steps: int, peak: float, flat_too: bool) -> list:
    """Measurements beside the headline (same JSON line, `secondary`): what happens when visibility is NOT per
    subtree (inside camera; a flat scene with per-entity random visibility), the update and cull passes on their
    own, and the expansion of mixed records.  Every entry: ms_per_step device-timed like the headline."""
    zn = rig.zn
    out = []

    def frame_entry(name, sc, n, n_roots, eye, note):
        sc.SetCamera(zn.Camera(zn.CameraDesc(eye=eye, center=(eye[0], eye[1], eye[2] - 1.0))))
        for _ in range(3):
            sc.Update()
        vis = sc.VisibleCount()
        ms = rig.timed(sc.Update, steps) / steps
        _, tails = profiled_levels(sc, min(steps, 10))
        alg = n_roots * 188 + (n - n_roots) * 192 + 4 * vis
        return {"name": name, "what": note, "entities": n, "camera_eye": list(eye), "visible_fraction": vis / n, "ms_per_step": ms,
                "value": n / (ms * 1e-3) / 1e6, "unit": UNIT, "frame_frac": alg / (ms * 1e-3) / 1e9 / peak, "emit_ms": float(tails.mean())}, vis

    mid_vis = None
    for name, eye, note in (("h8geo16m/inside", (0.0, 0.0, 0.0), "camera inside the scene: mixed masks in most chunks, the general compaction path"),):
        entry, vis = frame_entry(name, scene, E, roots, eye, note)
        entry["expand7"] = expand_probe(rig, scene, E, vis)
        out.append(entry)
    # back to the headline camera for the pass-level entries
    scene.SetCamera(zn.Camera(zn.CameraDesc(eye=CAMERA_EYE, center=(CAMERA_EYE[0], CAMERA_EYE[1], CAMERA_EYE[2] - 1.0))))
    scene.Update()
    mid_vis = scene.VisibleCount()
    cam = zn.Camera(zn.CameraDesc(eye=CAMERA_EYE, center=(CAMERA_EYE[0], CAMERA_EYE[1], CAMERA_EYE[2] - 1.0)))
    # update (world only) and cull as the two calls of the reference's frame loop
    for _ in range(3):
        scene.UpdateWorld()
    ms_world = rig.timed(scene.UpdateWorld, steps) / steps
    world_bytes = roots * 100 + (E - roots) * 104
    out.append({"name": "h8geo16m/update_world", "what": "zn_scene_update_world: TRS -> world over the hierarchy only (36 B in + 4 B parent + 64 B out per entity)",
                "entities": E, "ms_per_step": ms_world, "value": E / (ms_world * 1e-3) / 1e6, "unit": UNIT,
                "algorithmic_bytes": world_bytes, "frac": world_bytes / (ms_world * 1e-3) / 1e9 / peak})
    for _ in range(3):
        scene.Cull(cam)
    ms_cull = rig.timed(lambda: scene.Cull(cam), steps) / steps
    scene.SetProfiling(True)
    kt = []
    for _ in range(min(steps, 10)):
        scene.Cull(cam)
        kt.append(scene.CullTimesMs())
    scene.SetProfiling(False)
    kt = np.asarray(kt, np.float64)
    assert scene.VisibleCount() == mid_vis
    out.append({"name": "h8geo16m/cull_pass", "what": "zn_scene_cull: one camera pass over existing world matrices (64 + 24 B in, 64 B out per entity, + 4 B per visible index)",
                "entities": E, "visible_fraction": mid_vis / E, "ms_per_step": ms_cull, "value": E / (ms_cull * 1e-3) / 1e6, "unit": UNIT,
                "algorithmic_bytes": E * 152 + 4 * mid_vis, "frac": (E * 152 + 4 * mid_vis) / (ms_cull * 1e-3) / 1e9 / peak,
                "cull_kernel_ms": float(kt[:, 0].mean()), "cull_kernel_frac": E * 152 / (kt[:, 0].mean() * 1e-3) / 1e9 / peak,
                "emit_ms": float(kt[:, 1].mean())})
    out.append(dict(expand_probe(rig, scene, E, mid_vis), name="h8geo16m/mid/expand7",
                    what="expansion of seven copies of the headline frame's record (what a rank runs per frame at N = 8)"))
    if flat_too:
        flat = zn.Scene(rig.ctx, FLAT16M)
        flat.FillSynthetic(BASE_SEED + 7, 1, 0.0)
        entry, vis = frame_entry("flat16m/random", flat, FLAT16M, FLAT16M, CAMERA_EYE,
                                 "16,777,216 root entities at random positions: per-entity (incoherent) visibility, the worst case for the compaction")
        entry["expand7"] = expand_probe(rig, flat, FLAT16M, vis)
        out.append(entry)
        flat.Close()
    return out


def make_exchange(rig: Rig, mode: str, words: int, E: int, records_per_rank: int = 1):
    """(exchange, mode actually used, note).  `peer` = the C++ host layer (zn_exchange_*): torch only carries the 64-byte
    IPC handles at set-up.  The torch-hosted variants are kept for comparison."""
    torch = rig.torch
    note = ""
    if mode in ("peer", "peer-overlap"):
        from zenyth_b200.peer_exchange import PeerExchange
        try:
            return PeerExchange(rig.ctx, words, E, rig.world_size, rig.rank, records_per_rank=records_per_rank, slots=4,
                                overlap=mode == "peer-overlap"), mode, note
        except RuntimeError as e:       # raised on every rank together: peer mapping is not possible on this box
            note, mode = f"--exchange peer unavailable ({e}); ran --exchange records", "records"
    if mode.startswith("peer"):
        from zenyth_b200.peer_exchange import PeerRecordExchange
        try:
            return PeerRecordExchange(rig.ctx, words, E, rig.world_size, rig.rank, rig.dev, torch.int32, slots=4, records_per_rank=records_per_rank,
                                      signal="nccl" if mode == "peer-nccl" else "flags", fused_wait=mode != "peer-waitkernel"), mode, note
        except RuntimeError as e:
            note, mode = f"--exchange {mode} unavailable ({e}); ran --exchange records", "records"
    if mode == "records":
        from zenyth_b200.distributed import RecordExchange
        return RecordExchange(words, E, rig.world_size, rig.dev, torch.int32, slots=2, records_per_rank=records_per_rank, ctx=rig.ctx), mode, note
    from zenyth_b200.distributed import VisibleExchange
    return VisibleExchange(E, rig.world_size, rig.dev, torch.int32, slots=2, ctx=rig.ctx), "lists", note


COLLECTIVE_NOTES = {
    "lists": "NCCL all_gather of visible counts + index lists every step, double-buffered: the exchange of frame k overlaps the update of "
             "frame k+1; the timed region ends after the last exchange",
    "records": "NCCL all_gather of the fixed-size visibility records (per-chunk counts + 1 mask bit per entity) every step, double-buffered; "
               "every rank then expands all records into the ordered index lists + counts (identical to an all-gather of the lists, 27x "
               "fewer NVLink bytes); the timed region ends after the last expansion",
    "peer": "fused into the compute kernels over NVLink peer memory, no collective and no torch on the data path (C++ host layer zn_exchange_*): "
            "the level and emit kernels store the visibility record tile by tile into every peer's buffer (CUDA IPC mapped pointers); the kernel "
            "that follows the update on the stream raises a sequence flag at every peer (one fence.sys + relaxed sys-scope stores) and the "
            "expansion kernel's warps wait on the owners' flags themselves; every rank expands its peers' records into the ordered index "
            "lists + counts one frame behind (its own list comes straight from its emit kernel); the timed region ends after the last expansion",
    "peer-overlap": "as peer, but the expansion of frame k runs on a low-priority side stream right behind update(k) while the context stream "
                    "already runs update(k+1); update(k) waits for the rank's own expansion of frame k-2 (ZN_EXCHANGE_OVERLAP)",
    "peer-nccl": "peer stores of the records + a 4-byte NCCL all-reduce per frame as the cross-rank barrier (torch-hosted variant)",
    "peer-waitkernel": "peer stores of the records; signal + wait as a launch of their own (torch-hosted variant)",
    "none": "none (1 GPU)",
}


def checksum(torch, dev, t):
    t = t.to(torch.int64) & 0xFFFFFFFF
    return torch.stack([t.sum(), (t * (torch.arange(t.numel(), device=dev) % 1021 + 1)).sum()])


def run_ours(args) -> None:
    rig = Rig()
    torch, dist, zn, dev, stream, ctx = rig.torch, rig.dist, rig.zn, rig.dev, rig.stream, rig.ctx
    world_size, rank, distributed = rig.world_size, rig.rank, rig.distributed

    roots, levels, fanout, desc = WORKLOADS[args.workload]
    sizes = level_sizes(roots, levels, fanout)
    E = sum(sizes)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    scene = zn.Scene(ctx, E, offsets if levels > 1 else None)
    scene.FillSynthetic(BASE_SEED + 3 + rank, fanout, 0.0)
    cam = zn.Camera(zn.CameraDesc(eye=CAMERA_EYE, center=(CAMERA_EYE[0], CAMERA_EYE[1], CAMERA_EYE[2] - 1.0)))
    scene.SetCamera(cam)
    scene.SetIndexBase(rank * E if distributed and rank * E < 2 ** 32 else 0)
    bufs = scene.DeviceBuffers()
    launches_per_update = scene.LaunchesPerUpdate()

    frame_no = [0]
    index_bases = np.asarray([(r * E) % 2 ** 32 for r in range(world_size)], np.uint32)
    mode, fallback_note, exchange = "none", "", None
    if distributed:
        exchange, mode, fallback_note = make_exchange(rig, args.exchange, bufs.visibility_record_words, E)
    use_lists, use_records, use_peer = mode == "lists", mode == "records", mode.startswith("peer")

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
        elif mode == "peer-overlap":
            k = frame_no[0]
            exchange.bind(scene, k)                    # (update(k) waits for this rank's own expand(k-2): flow control)
            scene.Update()                             # kernels write this rank's record into every peer
            exchange.expand(scene, k, index_bases)     # side stream: raises frame k's flags, waits on the device for the peers', derives
                                                       # their lists while the next Update() already runs on the context stream
        elif use_peer:
            k = frame_no[0]
            exchange.bind(scene, k)
            scene.Update()                             # kernels write this rank's record into every peer
            exchange.launch(k)                         # (peer-nccl only: barrier(k), asynchronous)
            exchange.expand(scene, k - 1, index_bases)   # raises frame k's flags; waits on the device for the peers' frame k-1 flags
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
        if mode == "peer-overlap":
            exchange.join()                            # the context stream (and the timing event on it) waits for the last expansion
        elif use_peer:
            for k in (frame_no[0] - 2, frame_no[0] - 1):
                exchange.expand(scene, k, index_bases)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    finish()
    rig.barrier()
    if use_lists:
        scene.BindVisibleOutput(None)
        scene.Update()
    visible = scene.VisibleCount()
    all_visible = [visible]
    if use_records or use_peer:   # every rank derived the same lists: check them against each rank's own count
        counts, lists = exchange.result()
        own = torch.tensor([visible], dtype=torch.int32, device=dev)
        every = torch.zeros(world_size, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(every, own)
        assert counts == every.tolist(), (counts, every.tolist())
        all_visible = counts
        # ... and every derived list against its owner's actual list (checksums travel, not lists)
        mine = torch.from_numpy(scene.Visible().astype(np.int64)).to(dev)
        sums = torch.zeros(world_size * 2, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, checksum(torch, dev, mine))
        for r in range(world_size):
            assert torch.equal(checksum(torch, dev, lists[r, : counts[r]]), sums[2 * r: 2 * r + 2]), f"rank {rank}: list of rank {r} differs from its owner's"

    sampler = make_sampler(rig.gpu_id)
    total_ms = rig.timed(step, args.steps, finish)
    clocks = sampler.stop()

    # ---- sustained: seconds of back-to-back frames, clocks sampled throughout ----------------------
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / max(total_ms / args.steps, 1e-3)))
        sampler = make_sampler(rig.gpu_id)
        sus_ms = rig.timed(step, n_sus, finish)
        sus_clocks = sampler.stop()
        sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus,
                     "value": world_size * E * n_sus / (sus_ms * 1e-3) / 1e6, "unit": UNIT, "clocks": sus_clocks}
    if use_lists:
        scene.BindVisibleOutput(None)
    if use_records:
        scene.BindVisibilityRecord(None)
    if use_peer:
        exchange.close(scene)

    # ---- dominant kernel: last (largest) level, timed inside the library with CUDA events --------
    level_ms, tails = profiled_levels(scene, args.steps)
    dom = int(np.argmax(level_ms.mean(axis=0)))
    dom_ms = float(level_ms[:, dom].mean())
    dom_entities = int(sizes[dom])
    dom_bytes = dom_entities * (188 if dom == 0 else 192)
    peak, peak_src = measured_peaks()
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    frame_bytes = scene.AlgorithmicBytes(visible)
    frame_ms_prof = float(level_ms.sum(axis=1).mean() + tails.mean())
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
    scaling_ceiling = None
    if distributed:
        # every rank must also WRITE the N-1 peer lists: the byte ceiling of "every list on every rank" at this visible fraction
        others = sum(all_visible) - visible
        scaling_ceiling = frame_bytes / (frame_bytes + 4 * others)

    # ---- secondary measurements -------------------------------------------------------------------
    secondary = []
    if not args.no_secondary and args.workload == "h8geo16m":
        secondary = secondary_lines(rig, scene, E, roots, args.steps, peak, flat_too=True)
    scene.SetCamera(cam)

    # ---- e2e: the C-ABI call a CPU engine swaps in, HOST buffers, copies inside the timed region --
    e2e = None
    if not args.no_e2e:
        inputs = scene.Inputs()
        h_pos, h_rot, h_scl = (ctx.HostAlloc((E, 3), np.float32) for _ in range(3))
        h_pos[:] = inputs["position"]; h_rot[:] = inputs["rotation"]; h_scl[:] = inputs["scale"]
        del inputs
        h_vis = [ctx.HostAlloc((E,), np.uint32) for _ in range(2)]
        e2e_steps = args.steps
        barrier_host = dist.barrier if distributed else (lambda: None)
        # (1) one frame at a time: upload, update, download, return (zn_scene_frame_host)
        for _ in range(2):
            n_vis = scene.FrameHost(h_pos, h_rot, h_scl, visible=h_vis[0])
        rig.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            n_vis = scene.FrameHost(h_pos, h_rot, h_scl, visible=h_vis[0])
        torch.cuda.synchronize(dev)
        dt_sync = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)

        def pipelined(**arrays):
            """Two frames in flight (zn_scene_frame_host_submit / _wait): every step uploads its arrays and downloads its
            count + list; frame k+1's upload overlaps frame k's update and download.  Steady state, host wall clock."""
            scene.SubmitFrameHost(**arrays)
            scene.SubmitFrameHost(**arrays)
            for k in range(2):
                n = scene.WaitFrameHost(h_vis[k % 2])
                scene.SubmitFrameHost(**arrays)
            barrier_host()
            t0 = time.perf_counter()
            for k in range(e2e_steps):
                n = scene.WaitFrameHost(h_vis[k % 2])
                scene.SubmitFrameHost(**arrays)
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            for k in range(2):
                scene.WaitFrameHost(h_vis[k % 2])
            if distributed:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item()), n

        # (2) the full upload: position + rotation + scale, 36 B per entity
        dt_pipe, n_pipe = pipelined(position=h_pos, rotation=h_rot, scale=h_scl)
        assert n_pipe == n_vis, (n_pipe, n_vis)
        # (3) a frame that only spins its entities ("yaw += dt", the survey's config-1 pattern): rotation alone, 12 B per entity
        dt_rot, n_rot = pipelined(rotation=h_rot)
        assert n_rot == n_vis, (n_rot, n_vis)
        if distributed:
            dist.all_reduce(dt_sync, op=dist.ReduceOp.MAX)
        ceiling = h2d_ceiling(rig, int(E) * 36)
        h2d_rate = world_size * int(E) * 36 * e2e_steps / dt_pipe / 1e9
        e2e = {"value": world_size * E * e2e_steps / dt_pipe / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(E) * 36, "d2h_bytes_per_step": int(n_vis) * 4 + 4,
               "ms_per_step": dt_pipe / e2e_steps * 1e3,
               "api": "zn_scene_frame_host_submit/_wait, two frames in flight: every step uploads pinned host "
                      "position/rotation/scale and downloads the visible count + index list; world/MVP stay "
                      "device-resident for the renderer",
               "roofline": {"bound": "pcie", "achieved": h2d_rate, "peak": ceiling["aggregate_GBs"], "unit": "GB/s",
                            "frac": h2d_rate / ceiling["aggregate_GBs"], "per_gpu_peak": ceiling["per_gpu_GBs"],
                            "peak_source": ceiling["how"],
                            "note": "the step IS the 36 B/entity upload: achieved = aggregate H2D bytes per second of the e2e loop, peak = "
                                    "what a bare pinned cudaMemcpyAsync of the same size reaches on this box with all ranks copying at once"},
               "rotation_only": {"value": world_size * E * e2e_steps / dt_rot / 1e6, "unit": UNIT, "ms_per_step": dt_rot / e2e_steps * 1e3,
                                 "h2d_bytes_per_step": int(E) * 12, "d2h_bytes_per_step": int(n_vis) * 4 + 4,
                                 "api": "the same call with position = scale = NULL: a frame that only spins its entities uploads 12 B/entity"},
               "one_frame_at_a_time": {"value": world_size * E * e2e_steps / float(dt_sync.item()) / 1e6, "unit": UNIT,
                                       "ms_per_step": float(dt_sync.item()) / e2e_steps * 1e3, "api": "zn_scene_frame_host"},
               "numa": rig.numa, "timer": "host wall clock around the calls, steady state"}

    # ---- N > 1: BASELINE configs[4] (8 x 32M scenes, strong scaling) rides along as a secondary entry ----
    scene.Close()
    if distributed and not args.no_secondary and args.workload == "h8geo16m":
        try:
            secondary.append(dict(scenes8_measure(rig, args, args.exchange), name="scenes8x32m/strong"))
        except Exception as e:      # noqa: BLE001 - a secondary entry must not take the headline down
            secondary.append({"name": "scenes8x32m/strong", "error": f"{type(e).__name__}: {e}"})

    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.workload)

    if rank == 0:
        value = world_size * E * args.steps / (total_ms * 1e-3) / 1e6
        step_ms = total_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, E, levels, visible),
            "exchange": {"mode": mode, "collective": COLLECTIVE_NOTES[mode] + ((" [" + fallback_note + "]") if fallback_note else ""),
                         "regathers": getattr(exchange, "regathers", None) if use_lists else None},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "frac can exceed 1: the peak is torch's copy_ (1 read : 1 write); this kernel moves 1 read : 1.73 write "
                                 "in 16 KiB TMA bursts and reaches 83% of the 8.18 TB/s ncu reports as the device's DRAM peak "
                                 "(profiles/r01_s2_level7_ncu_summary.csv), against 79% for the copy",
                         "kernel": f"scene_level_kernel<{'true' if dom else 'false'}> level {dom} ({dom_entities:,} entities)",
                         "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": dom_ms,
                         "scaling_ceiling": scaling_ceiling,
                         "frame": {"algorithmic_bytes": frame_bytes, "ms": step_ms,
                                   "achieved": frame_bytes / (step_ms * 1e-3) / 1e9,
                                   "frac": frame_bytes / (step_ms * 1e-3) / 1e9 / peak,
                                   "level_ms": [float(x) for x in level_ms.mean(axis=0)], "scan_emit_ms": float(tails.mean()),
                                   "profiled_frame_ms": frame_ms_prof}},
            "sustained": sustained,
            "secondary": secondary,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": args.steps * (launches_per_update + (1 if (use_records or use_peer) else 0)),
        }
        emit_line(json.dumps(line), flush=True)
    rig.close()


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


def scenes8_measure(rig: Rig, args, exchange_mode: str) -> dict:
    """BASELINE configs[4]: 8 independent 33,554,430-entity H8-geo scenes sharded over N = 1/2/4/8
    GPUs (8/N scenes per GPU, updated back to back), then the exchange: every rank ends up with all
    8 compacted visible lists and counts.  Strong scaling: the total work is fixed."""
    from zenyth_b200.distributed import scene_index_bases, shard_scenes
    torch, dist, zn, dev, ctx = rig.torch, rig.dist, rig.zn, rig.dev, rig.ctx
    world_size, rank, distributed = rig.world_size, rig.rank, rig.distributed
    total_scenes = 8
    assert total_scenes % world_size == 0
    owned = shard_scenes(total_scenes, world_size, rank)
    sizes = level_sizes(14, 8, 8)
    E = sum(sizes)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    bases = scene_index_bases([E] * total_scenes)
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
    mode, fallback_note, exchange = "none", "", None
    if distributed:
        if exchange_mode == "lists":
            exchange_mode = "peer"          # this workload exchanges visibility records
        exchange, mode, fallback_note = make_exchange(rig, exchange_mode, words, E, records_per_rank=len(owned))
    use_peer = mode.startswith("peer")
    index_bases = np.asarray(bases, np.uint32)

    def step():
        k = frame_no[0]
        if mode == "peer-overlap":
            exchange.bind(scenes, k)
            for sc in scenes:
                sc.Update()                              # kernels store the records into every peer as they run
            exchange.expand(scenes[0], k, index_bases)   # side stream, overlaps the next step's updates
        elif use_peer:
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
        if mode == "peer-overlap":
            exchange.join()
        elif use_peer:
            for k in (frame_no[0] - 2, frame_no[0] - 1):
                exchange.expand(scenes[0], k, index_bases)
        elif distributed:
            for slot in ((frame_no[0] - 2) % 2, (frame_no[0] - 1) % 2):
                if exchange.ready(slot):
                    exchange.expand(scenes[0], slot, index_bases)

    for _ in range(max(args.warmup, 3)):
        step()
    finish()
    rig.barrier()
    local_counts = [sc.VisibleCount() for sc in scenes]
    if distributed:
        # every rank derived all 8 lists: check counts and list checksums against the owners' own lists
        counts, lists = exchange.result()
        own = torch.stack([torch.cat([torch.tensor([n], dtype=torch.int64, device=dev),
                                      checksum(torch, dev, torch.from_numpy(sc.Visible().astype(np.int64)).to(dev))])
                           for sc, n in zip(scenes, local_counts)]).flatten()
        every = torch.zeros(total_scenes * 3, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, own)
        every = every.view(total_scenes, 3)
        assert counts == every[:, 0].tolist(), (counts, every[:, 0].tolist())
        for sid in range(total_scenes):
            assert torch.equal(checksum(torch, dev, lists[sid, : counts[sid]]), every[sid, 1:]), f"rank {rank}: list of scene {sid} differs from its owner's"
    else:
        counts = local_counts
    steps = max(3, min(args.steps, 20))
    total_ms = rig.timed(step, steps, finish)
    if use_peer:
        exchange.close(scenes)
    for sc in scenes:
        if distributed and not use_peer:
            sc.BindVisibilityRecord(None)
        sc.Close()
    peak, peak_src = measured_peaks()
    alg = total_scenes * (14 * 188 + (E - 14) * 192) + 4 * sum(counts)
    step_ms = total_ms / steps
    own_bytes = len(owned) * (14 * 188 + (E - 14) * 192) + 4 * sum(local_counts)
    others = sum(counts) - sum(local_counts)
    return {"metric": METRIC, "value": total_scenes * E * steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world_size,
            "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "8 independent H8-geo scenes of 33,554,430 entities sharded over the GPUs (BASELINE configs[4])",
                       "scenes_per_gpu": len(owned), "visible_counts": counts, "l2": l2_note(E)},
            "exchange": {"mode": mode, "collective": ("none (1 GPU: every scene's own emit already leaves its list on the device)" if not distributed
                                                      else COLLECTIVE_NOTES[mode]) + ((" [" + fallback_note + "]") if fallback_note else "")},
            "roofline": {"bound": "hbm", "achieved": alg / (step_ms * 1e-3) / 1e9 / world_size, "peak": peak, "unit": "GB/s",
                         "frac": alg / (step_ms * 1e-3) / 1e9 / world_size / peak, "traffic": None, "peak_source": peak_src,
                         "scaling_ceiling": own_bytes / (own_bytes + 4 * others) if distributed else None,
                         "kernel": "whole step, per GPU (algorithmic bytes of all 8 frames / step time / N)"},
            "gpu_launches": steps * (len(owned) * launches_per_update + (1 if distributed else 0))}


def run_scenes8(args) -> None:
    rig = Rig()
    line = scenes8_measure(rig, args, args.exchange)
    if rig.rank == 0:
        emit_line(json.dumps(line), flush=True)
    rig.close()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["mesh64m", "scenes8x32m"], default="h8geo16m")
    ap.add_argument("--exchange", choices=["peer", "peer-overlap", "records", "lists", "peer-nccl", "peer-waitkernel"], default="peer",
                    help="N>1: the kernels store the visibility records straight into peer GPUs' memory over NVLink and signal "
                         "completion with flags in peer memory (peer, default; falls back to records where peer mapping is refused); "
                         "NCCL all-gather of the records (records) or of the index lists (lists); peer stores with a 4-byte "
                         "all-reduce as the barrier (peer-nccl) or with signal + wait as a kernel of their own (peer-waitkernel)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` measurements (inside camera, flat 16M, update/cull passes, expansion; N>1: scenes8x32m)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the `sustained` sub-record")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained back-to-back run")
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
