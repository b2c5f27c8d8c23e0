#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 raster path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode frames|tiles]

One "step" = one whole frame of the hot path (geometry -> binning -> tile raster -> resolve) of the
1M-triangle textured UV-sphere at 1920x1080, 4xMSAA, bilinear filtering (BASELINE.json configs[1]).
At N>1 the ranks render independent frames of the orbiting-camera sweep of the same mesh
(configs[4]); there is no data-path collective (DESIGN.md "Multi-GPU"), so scaling is "weak".
`--mode tiles` instead splits ONE 8192x8192 frame into tile-row ranges per rank and gathers the
strips with NCCL (configs[3]).

Prints ONE JSON line on rank 0 (contract in the task statement):
  value      Mtris/s with mesh, texture and uniforms already resident in HBM (CUDA events)
  e2e        same metric through the host-buffer API: H2D of the mesh + D2H of the image every step
  roofline   dominant kernel: algorithmic bytes / measured kernel time vs the measured HBM peak
  cpu_baseline  the CPU oracle (C port of the reference algorithm) timed on this host, 1 core
  cpu_baseline_c1  the same for BASELINE configs[0] (the crate's default 1280x720 scene, the config named "on CPU")
  tiles      (N > 1) BASELINE configs[3]: one 8192x8192 frame split into tile rows across the N GPUs, timed with the
             NCCL gather and with NVLink peer stores, checked against the single-GPU frame
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "Mtris/s at 1080p 4xMSAA (1M-triangle textured frame)"
UNIT = "Mtris/s"
HBM_FALLBACK_GBS = 6650.0


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every few
    milliseconds (the timed region of the default run lasts tens of milliseconds); `nvidia-smi -lms` as fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, period_s: float = 0.0005):
        self.index, self.period, self.rows, self.proc, self.thread, self.stop_flag = index, period_s, [], None, None, False
        self.mode = None

    def _reasons(self, mask):
        import pynvml as N

        bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
        return [n for n, b in bits.items() if mask & b]

    def sample_now(self):
        """One NVML reading right now (called immediately before and after the timed block, so even a timed region of
        a millisecond carries clock evidence)."""
        if self.mode != "nvml":
            return
        import pynvml as N

        try:
            sm = N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)
            mask = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            self.rows.append((time.time(), float(sm), float(self.max_sm), self._reasons(mask)))
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag:
            self.sample_now()
            time.sleep(self.period)

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            # NVML enumerates all GPUs of the box: honour CUDA_VISIBLE_DEVICES when it lists ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            self.h = N.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.mode = "nvidia-smi"
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.time(), float(f[1]), float(f[2]),
                                  [n for n, v in zip(self.NAMES, f[5:9]) if v.lower().startswith("active")]))
            except Exception:
                continue

    def stop(self, t0, t1):
        if self.mode is None:
            return None
        if self.mode == "nvml":
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            margin0, margin1 = 0.0, 0.0
        else:
            time.sleep(0.15)
            self.proc.terminate()
            margin0, margin1 = 0.05, 0.15
        rows = [r for r in self.rows if t0 - margin0 <= r[0] <= t1 + margin1]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.mode}
        reasons = sorted({n for r in rows for n in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows), "reasons": reasons,
                "samples": len(rows), "source": self.mode}


def bind_host_near_gpu(local: int, world: int):
    """Pin this rank to host cores near its GPU (NVML's ideal CPU affinity = the GPU's NUMA node) before any pinned
    buffer is allocated, so first-touch places the staging memory on that node.  When several ranks share one affinity
    set (one NUMA node for all GPUs) the set is dealt out evenly among them.  Returns what was done (for the JSON line)."""
    try:
        import pynvml as N

        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        h = N.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = N.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = sorted(c for c in range(ncpu) if (words[c // 64] >> (c % 64)) & 1)
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in ideal if c in allowed] or allowed
        if world > 1 and len(cpus) >= world:  # ranks that share the set take disjoint slices of it
            per = len(cpus) // world
            cpus = cpus[local * per:(local + 1) * per]
        os.sched_setaffinity(0, cpus)
        numa = None
        try:
            bus = N.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            numa = int(open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read())
        except Exception:
            pass
        return {"cpus": f"{cpus[0]}-{cpus[-1]}" if cpus else None, "n_cpus": len(cpus), "gpu_ideal_cpus": len(ideal), "gpu_numa_node": numa}
    except Exception as e:  # no NVML / not permitted: run unbound
        return {"error": str(e)[:80]}


def tiles_subrecord(args, rank, world, local):
    """BASELINE configs[3] inside the N-GPU run: ONE 8192x8192 frame of the 1M-triangle sphere, geometry replicated, every
    rank rasterising its tile rows; timed (a) with contiguous row ranges + one NCCL all_gather of the strips and (b) with
    interleaved bands of tile rows whose tile kernels store straight into rank 0's image over NVLink (CUDA IPC peer
    memory, flag kernels; sharding.PeerFrame).  Rank 0 also renders the frame alone: that is the strong-scaling
    baseline and the image the assembled frames must equal.  Per-frame CUDA events, L2 flushed between frames,
    median over frames, max over ranks."""
    import torch
    import torch.distributed as dist

    from rusterizer_b200 import scenes
    from rusterizer_b200.render import Renderer
    from rusterizer_b200.sharding import PeerFrame

    W = H = 8192
    scene = scenes.sphere_scene(args.n_phi, args.n_theta, width=W, height=H)
    mesh = scene.draws[0].mesh
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K, Wm = 12, 3

    def make_renderer():
        r = Renderer(W, H, device=local)
        r.set_stream(stream.cuda_stream)
        r.uniforms().bind_texture(0, scene.texture)
        dm = r.upload(mesh)
        blk = r.uniforms().write_block()
        blk.projection, blk.world, blk.view = scene.projection, scene.draws[0].world, scene.view
        return r, dm

    def timed(frame, r):
        for _ in range(Wm):
            frame()
        r.sync()
        torch.cuda.synchronize()
        dist.barrier()
        ev = []
        for _ in range(K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            frame()
            b.record(stream)
            ev.append((a, b))
        r.sync()
        torch.cuda.synchronize()
        t = torch.tensor([statistics.median(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class _Raw:  # zero-copy torch view of a device pointer owned by the library
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (H, W), "typestr": "<i4", "data": (int(ptr), False), "version": 3}

    solo_ms, want = None, None
    if rank == 0:  # the whole frame on one GPU
        r, dm = make_renderer()
        r.render(dm, 0, 0)
        r.framebuffer_device()
        ts = []
        for i in range(Wm + K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            r.render(dm, 0, 0)
            r.framebuffer_async()
            b.record(stream)
            r.sync()
            ts.append(a.elapsed_time(b))
        solo_ms = statistics.median(ts[Wm:])
        r.render(dm, 0, 0)
        want = torch.as_tensor(_Raw(r.framebuffer_device()), device="cuda").clone()
        r.close()
    dist.barrier()
    th = 16
    rows_per = ((H // th + world - 1) // world) * th
    r0, r1 = min(H, rank * rows_per), min(H, (rank + 1) * rows_per)
    out = {"workload": "BASELINE configs[3]: 8192x8192, 1M-triangle sphere, tile rows split over %d GPUs" % world,
           "single_gpu_ms_per_frame": solo_ms, "root_ingest_bytes_per_frame": (world - 1) * W * H * 4 // world,
           "nvlink_peak_gb_per_s_one_way": 900.0, "frames_timed": K, "l2": "flushed between frames (256 MiB memset, not timed)"}
    for mode in ("nccl", "peer"):
        r, dm = make_renderer()
        pf = None
        if mode == "nccl":
            r.set_row_range(r0, r1)
            strip = torch.empty((rows_per, W), dtype=torch.int32, device="cuda")
            gather = torch.empty((world * rows_per, W), dtype=torch.int32, device="cuda")
            last = [None]

            def frame():
                r.render(dm, 0, 0)
                r.framebuffer_async(strip.data_ptr())
                dist.all_gather_into_tensor(gather, strip)
                last[0] = gather[:H]
        else:
            pf = PeerFrame(r, root=0, n_buffers=2, interleave_band=args.band)
            last = [None]

            def frame():
                r.render(dm, 0, 0)
                img = pf.finish_frame()
                pf.release()
                if img is not None:
                    last[0] = torch.as_tensor(_Raw(img), device="cuda")
        r.render(dm, 0, 0)
        r.framebuffer_device()  # the synchronous call sizes the device buffers
        ms = timed(frame, r)
        frame()
        r.sync()
        torch.cuda.synchronize()
        dist.barrier()
        ok = torch.tensor([1 if (rank != 0 or torch.equal(last[0], want)) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dist.barrier()
        key = "nccl_gather" if mode == "nccl" else "peer_stores"
        out[key] = {"ms_per_frame": ms, "assembled_frame_matches_single_gpu": bool(ok.item()),
                    "rows": ("contiguous tile-row ranges" if mode == "nccl" else f"interleaved bands of {args.band} tile rows"),
                    "strong_scaling_efficiency": (None if solo_ms is None else solo_ms / (world * ms)),
                    "root_ingest_gb_per_s": (world - 1) * W * H * 4 / world / (ms / 1e3) / 1e9}
        if pf is not None:
            pf.close()
        r.close()
        del last
    if solo_ms is not None:
        for k in ("nccl_gather", "peer_stores"):
            out[k]["speedup_vs_single_gpu"] = solo_ms / out[k]["ms_per_frame"]
    return out


def build_scene(args):
    from rusterizer_b200 import scenes

    if args.mode == "tiles":
        return scenes.sphere_scene(args.n_phi, args.n_theta, width=8192, height=8192)
    return scenes.sphere_scene(args.n_phi, args.n_theta, width=args.width, height=args.height)


def cpu_baseline(scene, budget_s=12.0, fast=True):
    """Time the CPU oracle (the reference algorithm restated in C, single thread like the reference)
    on whole frames of the same workload until ~budget_s of CPU time is spent."""
    from oracle import oracle as orc

    try:
        lib = orc.OracleLib(orc.build(fast=fast, out_dir=Path(os.environ.get("TMPDIR", "/tmp")) / "rz_oracle_native"),
                            fast=fast)
        build = "gcc -O3 -march=native -ffp-contract=off"
    except Exception:
        lib = orc.get_lib(False)
        build = "gcc -O2 -ffp-contract=off"
    r = orc.OracleRenderer(scene.width, scene.height, lib)
    r.bind_texture(0, scene.texture.texels)
    r.write_block(view=scene.view, projection=scene.projection)

    def frame():
        for d in scene.draws:
            r.write_block(world=d.world)
            r.render(d.mesh.vertices, d.mesh.attributes, d.mesh.indices, 0, d.fs)
        return r.framebuffer()

    t = time.perf_counter()
    frame()  # warm-up
    first = time.perf_counter() - t
    times = []
    while sum(times) < budget_s and len(times) < 50:
        t = time.perf_counter()
        frame()
        times.append(time.perf_counter() - t)
        if first > budget_s:
            break
    best = min(times) if times else first
    cnt = r.counters()
    r.close()
    frames = len(times) + 1
    return {
        "value": scene.n_triangles / best / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"{frames} whole frames of the same workload ({scene.n_triangles} triangles, {scene.width}x{scene.height}), "
                  f"best frame {best * 1e3:.1f} ms; oracle/rz_oracle.c built with {build}; host has {os.cpu_count()} cores, "
                  "the reference is single-threaded",
        "ms_per_frame": best * 1e3,
        "gsamples_per_s": cnt["n_samples_written"] / frames / best / 1e9,
    }


def run_reference(args):
    """--impl reference: the reference's CPU path.  The crate is Rust and no Rust toolchain exists in
    this image, so the arm times the oracle port of the same algorithm (1 thread: the reference has
    no threading).  Each step is one whole frame of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = build_scene(args)
    steps = max(1, min(args.steps, 10))
    cb = cpu_baseline(scene, budget_s=min(60.0, 1.5 * steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_frame"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, scene),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gsamples_per_s": cb["gsamples_per_s"],
    }
    print(json.dumps(line), flush=True)


def workload_config(args, scene):
    return {
        "workload": ("BASELINE configs[3]: one 8192x8192 frame of the 1M-triangle sphere split into tile-row ranges + NCCL gather"
                     if args.mode == "tiles" else
                     "BASELINE configs[1]: 1M-triangle textured UV-sphere (1001x501), 1920x1080, 4xMSAA, bilinear checkerboard texture"
                     + ("; N>1: independent frames of the orbit sweep per GPU (configs[4])" if args.gpus > 1 else "")),
        "triangles": scene.n_triangles, "vertices": scene.n_vertices, "width": scene.width, "height": scene.height,
        "msaa": 4, "fs": "Texture", "parallelism": f"{args.mode}x{args.gpus}",
        "frames_in_flight_per_gpu": 1 if args.mode == "tiles" else max(1, args.inflight),
        "l2": ("flushed between timed frames (256 MiB memset outside the per-frame event pairs)" if args.mode == "tiles" else
               "inputs larger than L2: every context cycles through device copies of the mesh (>= 6 x 30 MB = 180 MB per GPU "
               "between two uses of a copy, L2 is 126 MB); one event pair around all timed frames, no flush kernel inside"),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="frames", choices=["frames", "tiles"])
    ap.add_argument("--n-phi", type=int, default=1001)
    ap.add_argument("--n-theta", type=int, default=501)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--inflight", type=int, default=6,
                    help="frames in flight per GPU in the timed region (one renderer context + stream each); frames mode only")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="tiles mode, N>1: 'peer' = tile kernels store their rows straight into rank 0's image over NVLink "
                         "(CUDA IPC peer memory + flag kernels); 'nccl' = local strips + one all_gather")
    ap.add_argument("--band", type=int, default=16,
                    help="tiles mode with --gather peer: tile rows per interleaved band (0 = contiguous row ranges)")
    ap.add_argument("--min-timed-s", type=float, default=0.25, help="repeat the K-step block until this much device time is timed")
    ap.add_argument("--max-blocks", type=int, default=400)
    ap.add_argument("--no-tiles", action="store_true", help="N > 1: skip the configs[3] sub-record (8192x8192 tile-row split)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from rusterizer_b200 import scenes
    from rusterizer_b200.camera import Camera
    from rusterizer_b200.render import Renderer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    host_binding = bind_host_near_gpu(local, world)  # before any pinned allocation
    if world > 1:
        # a failed or hung collective must surface as an error: the NCCL watchdog polls ncclCommGetAsyncError and tears the
        # communicator down instead of letting the ranks spin ("NCCL async error query")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    scene = build_scene(args)
    mesh = scene.draws[0].mesh
    W, H = scene.width, scene.height
    K, Wm = args.steps, args.warmup

    r = Renderer(W, H, device=local)
    # a real (non-default) torch stream: torch.cuda.Event times exactly the stream our kernels run on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    r.set_stream(stream.cuda_stream)
    r.uniforms().bind_texture(0, scene.texture)
    dmesh = r.upload(mesh)
    blk = r.uniforms().write_block()
    blk.projection = scene.projection
    blk.world = scene.draws[0].world

    tiles_mode = args.mode == "tiles"
    gather_buf, pf = None, None
    own_rows = H
    if tiles_mode:
        th = 16
        rows_per = ((H // th + n_gpus - 1) // n_gpus) * th
        r0, r1 = min(H, rank * rows_per), min(H, (rank + 1) * rows_per)
        r.set_row_range(r0, r1)
        own_rows = r1 - r0
        strip = torch.empty((rows_per, W), dtype=torch.int32, device="cuda")
        use_peer = n_gpus > 1 and args.gather == "peer"
        gather_buf = torch.empty((n_gpus * rows_per, W), dtype=torch.int32, device="cuda") if n_gpus > 1 and not use_peer else None
        if use_peer:
            from rusterizer_b200.sharding import PeerFrame

            pf = PeerFrame(r, root=0, n_buffers=2, interleave_band=args.band)
            own_rows = rows_per  # interleaved bands: the same number of rows per rank
        cams = [Camera()] * (K + Wm)
    elif n_gpus > 1:
        sweep = scenes.orbit_cameras(1024)
        cams = [sweep[(s * n_gpus + rank) % 1024] for s in range(K + Wm)]
    else:
        cams = [Camera()] * (K + Wm)
    views = [c.get_view_matrix() for c in cams]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def frame_async(i):
        blk.view = views[i]
        r.render(dmesh, 0, 0)
        if tiles_mode and pf is not None:
            last_image[0] = pf.finish_frame()  # rank 0: the complete frame, stream-ordered
            pf.release()
        elif tiles_mode:
            r.framebuffer_async(strip.data_ptr())
            if gather_buf is not None:
                dist.all_gather_into_tensor(gather_buf, strip)
        else:
            r.framebuffer_async()

    last_image = [None]
    # ---- warm-up: the synchronous path sizes the device buffers, then a few async frames ----
    blk.view = views[0]
    r.render(dmesh, 0, 0)
    r.framebuffer_device()
    for i in range(Wm):
        frame_async(i)
    r.sync()
    r.reset_counters()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- timed region: K frames, inputs resident in HBM ----
    # frames mode: `inflight` renderer contexts (own stream, own device buffers) take the frames round-robin, so the
    # geometry stage of one frame fills the SMs the tile stage of the previous frame leaves idle in its tail.  Every
    # context cycles through device copies of the mesh so that the inputs touched between two uses of a copy exceed
    # L2 (no flush kernel inside the timed region); ONE event pair brackets all K frames.
    # tiles mode: one context, L2 flushed between frames, per-frame event pairs (the flush is not timed).
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.02 if sampler.mode == "nvml" else 0.3)
    L = 1 if tiles_mode else max(1, args.inflight)
    lanes = [(r, stream, blk, [dmesh])]
    frame_latency_ms = None
    if not tiles_mode:
        copies = (6 + L - 1) // L  # >= 6 copies x 30 MB = 180 MB > 126 MB of L2
        for j in range(L):
            if j > 0:
                rj = Renderer(W, H, device=local)
                sj = torch.cuda.Stream()
                rj.set_stream(sj.cuda_stream)
                rj.uniforms().bind_texture(0, scene.texture)
                bj = rj.uniforms().write_block()
                bj.projection = scene.projection
                bj.world = scene.draws[0].world
                lanes.append((rj, sj, bj, []))
            rj, sj, bj, mj = lanes[j]
            while len(mj) < copies:
                mj.append(rj.upload(mesh))
            # size the device buffers through the synchronous path (it grows and replays) on a spread of the views this
            # lane will render: the async frames of the timed region must never outgrow them
            for v in sorted({0, len(views) // 7, 2 * len(views) // 7, 3 * len(views) // 7, 4 * len(views) // 7, 5 * len(views) // 7,
                             6 * len(views) // 7, len(views) - 1} if n_gpus > 1 else {0}):
                bj.view = views[v]
                rj.render(mj[0], 0, 0)
                rj.framebuffer_device()
            for w in range(max(Wm, len(mj))):
                bj.view = views[w % len(views)]
                rj.render(mj[w % len(mj)], 0, 0)
                rj.framebuffer_async()
            rj.sync()
            rj.reset_counters()
        # latency of one frame alone (one context, L2 flushed before it, per-frame event pairs)
        lat = []
        for s in range(min(K, 20)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            frame_async(Wm + s)
            b.record(stream)
            r.sync()
            lat.append(a.elapsed_time(b))
        frame_latency_ms = sum(lat) / len(lat)
        r.reset_counters()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    if world > 1:  # rank 0 slept while the sampler started: line the ranks up again right before the timed frames
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    launches0 = sum(x[0].launch_count() for x in lanes)
    if rank == 0:
        sampler.sample_now()
    t_wall0 = time.time()
    blocks, block_ev = 1, []
    if tiles_mode:
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for s in range(K):
            flush.zero_()
            ev0[s].record(stream)
            frame_async(Wm + s)
            ev1[s].record(stream)
        r.sync()
    else:
        # The K-step block is repeated back to back until >= args.min_timed_s of device time has been timed (a 20-step
        # block lasts 2 ms: too short for clock sampling and for a stable number).  ONE event pair brackets all blocks
        # (value = frames / that time); an event after every block gives the spread.  Every rank runs the same number
        # of blocks: it is derived from the warm-up estimate of rank 0.
        est = torch.tensor([max(frame_latency_ms or 0.1, 0.02) * 0.85 * K], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.broadcast(est, 0)
        blocks = int(min(args.max_blocks, max(1, -(-args.min_timed_s * 1e3 // float(est.item())))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step = 0
        for b in range(blocks):
            for s in range(K):
                rj, sj, bj, mj = lanes[step % L]
                bj.view = views[Wm + s]
                rj.render(mj[(step // L) % len(mj)], 0, 0)
                rj.framebuffer_async()
                step += 1
            for rj, sj, bj, mj in lanes[1:]:
                done = torch.cuda.Event()
                done.record(sj)
                stream.wait_event(done)
            eb = e1 if b == blocks - 1 else torch.cuda.Event(enable_timing=True)
            eb.record(stream)
            block_ev.append(eb)
        for x in lanes:
            x[0].sync()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.sample_now()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = sum(x[0].launch_count() for x in lanes) - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) if tiles_mode else e0.elapsed_time(e1)
    cnt = {}
    for x in lanes:
        for k, v in x[0].counters().items():
            cnt[k] = cnt.get(k, 0) + v
    for x in lanes[1:]:
        x[0].close()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    tot = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms_max = float(tot.item())
    frames_total = K * blocks * (1 if tiles_mode else n_gpus)
    value = (scene.n_triangles * frames_total) / (total_ms_max / 1e3) / 1e6
    block_ms = []
    if block_ev:
        prev = e0
        for eb in block_ev:
            block_ms.append(prev.elapsed_time(eb))
            prev = eb
    samples = torch.tensor([cnt["n_samples_written"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(samples, op=dist.ReduceOp.SUM)
    gsamples = float(samples.item()) / (total_ms_max / 1e3) / 1e9

    # ---- per-kernel times (CUDA events on the same stream) over K profiled frames ----
    stage = {"geometry_ms": [], "bin_ms": [], "tile_ms": [], "total_ms": []}
    for s in range(min(K, 20)):
        flush.zero_()
        blk.view = views[Wm + s]
        r.render(dmesh, 0, 0)
        r.framebuffer_device()
        t = r.timings()
        for k in stage:
            stage[k].append(t[k])
    stage_avg = {k: sum(v) / len(v) for k, v in stage.items()}

    # ---- tiles mode: the assembled frame on rank 0 must be the frame one GPU renders alone ----
    assembled_ok = None
    if tiles_mode and n_gpus > 1:
        frame_async(0)
        r.sync()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            if pf is not None:
                class _Raw:  # zero-copy torch view of the shared image (device pointer owned by the library)
                    __cuda_array_interface__ = {"shape": (H, W), "typestr": "<i4", "data": (int(last_image[0]), False), "version": 3}

                got = torch.as_tensor(_Raw(), device="cuda")
            else:
                got = gather_buf[:H]
            solo = Renderer(W, H, device=local)
            solo.uniforms().bind_texture(0, scene.texture)
            sb = solo.uniforms().write_block()
            sb.projection, sb.world, sb.view = scene.projection, scene.draws[0].world, views[0]
            solo.render(mesh, 0, 0)
            want = torch.from_numpy(solo.framebuffer().view(np.int32)).cuda()
            solo.close()
            assembled_ok = bool(torch.equal(got, want))
            if not assembled_ok:
                raise SystemExit("bench.py: the frame assembled from the ranks' tile rows differs from the single-GPU frame")
        dist.barrier()
    if pf is not None:
        pf.close()
        r.set_row_range(r0, r1)  # the e2e leg below runs each rank's contiguous strip

    # ---- e2e: host buffers in, host image out, every step (pinned memory) ----
    pos_h = torch.from_numpy(mesh.vertices).pin_memory()
    att_h = torch.from_numpy(mesh.attributes).pin_memory()
    idx_h = torch.from_numpy(mesh.indices.view(np.int32)).pin_memory()
    out_h = [torch.empty((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    e2e_steps = max(3, min(K, 30))

    def frame_e2e_sync(i):
        blk.view = views[i % len(views)]
        r.render_arrays(pos_h.data_ptr(), att_h.data_ptr(), mesh.n_vertices, idx_h.data_ptr(), mesh.indices.size, 0, 0)
        r.framebuffer_into(out_h[0].data_ptr())  # synchronises; D2H of the resolved image

    def frame_e2e(i):
        # the streaming form of render() + display(): every step uploads the host mesh (H2D), runs the frame
        # and reads the image back (D2H); copies of neighbouring steps overlap the kernels (rz.h)
        blk.view = views[i % len(views)]
        r.render_arrays(pos_h.data_ptr(), att_h.data_ptr(), mesh.n_vertices, idx_h.data_ptr(), mesh.indices.size, 0, 0)
        r.framebuffer_host_async(out_h[i % 2].data_ptr())

    # one-step latency of the synchronous call (no overlap), for reference
    for i in range(2):
        frame_e2e_sync(i)
    t0 = time.perf_counter()
    for i in range(5):
        frame_e2e_sync(i)
    e2e_sync_ms = (time.perf_counter() - t0) / 5 * 1e3
    ref_img = out_h[0].clone()  # view index 4

    for i in range(3):
        frame_e2e(i)
    r.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        frame_e2e(i)
    r.sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the streamed images are the ones the synchronous call returns (same view -> same bits)
    frame_e2e(4)
    r.sync()
    if not torch.equal(out_h[0], ref_img):
        raise SystemExit("bench.py: streamed e2e frame differs from the synchronous frame")
    # the same loop for a caller that keeps the mesh on the device (rz_mesh_upload once): per step only the uniform block
    # goes up and the image comes down.  The reference's mesh lives next to its renderer too; reported beside the
    # headline e2e figure, which pays the full mesh upload every step.
    def frame_e2e_resident(i):
        blk.view = views[i % len(views)]
        r.render(dmesh, 0, 0)
        r.framebuffer_host_async(out_h[i % 2].data_ptr())
    for i in range(3):
        frame_e2e_resident(i)
    r.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        frame_e2e_resident(i)
    r.sync()
    torch.cuda.synchronize()
    e2r = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    frame_e2e_resident(4)
    r.sync()
    if not torch.equal(out_h[0], ref_img):
        raise SystemExit("bench.py: resident-mesh e2e frame differs from the synchronous frame")
    e2 = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2r, op=dist.ReduceOp.MAX)
    e2e_resident = {"value": scene.n_triangles * e2e_steps * (1 if tiles_mode else n_gpus) / float(e2r.item()) / 1e6, "unit": UNIT,
                    "ms_per_step": float(e2r.item()) / e2e_steps * 1e3, "h2d_bytes_per_step": 192, "d2h_bytes_per_step": W * H * 4,
                    "how": "mesh uploaded once (rz_mesh_upload), per step: uniform block up, frame, image down "
                           "(rz_framebuffer_host_async); not the headline e2e"}
    e2e_value = scene.n_triangles * e2e_steps * (1 if tiles_mode else n_gpus) / float(e2.item()) / 1e6
    # PCIe / host-memory ceiling of this leg: the same bytes per step (mesh H2D on one stream, image D2H on another,
    # all ranks at once) with no kernels at all.  e2e ms / ceiling ms says how much of the e2e step is the copies.
    dev_in = torch.empty(pos_h.numel() * 4 + att_h.numel() * 4 + idx_h.numel() * 4, dtype=torch.uint8, device="cuda")
    dev_out = torch.empty((H, W), dtype=torch.int32, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    host_in = [pos_h.view(torch.uint8).reshape(-1), att_h.view(torch.uint8).reshape(-1), idx_h.view(torch.uint8).reshape(-1)]

    def copies(nsteps):
        for i in range(nsteps):
            with torch.cuda.stream(s_up):
                o = 0
                for h in host_in:
                    dev_in[o:o + h.numel()].copy_(h, non_blocking=True)
                    o += h.numel()
            with torch.cuda.stream(s_dn):
                out_h[i % 2].copy_(dev_out, non_blocking=True)
    copies(3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    copies(e2e_steps)
    torch.cuda.synchronize()
    cp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cp, op=dist.ReduceOp.MAX)
    copy_ms = float(cp.item()) / e2e_steps * 1e3
    del dev_in, dev_out
    h2d = mesh.vertices.nbytes + mesh.attributes.nbytes + mesh.indices.nbytes + 192
    d2h = W * H * 4

    # ---- N > 1: BASELINE configs[3] (one 8192x8192 frame split into tile rows) as a sub-record of the same line ----
    tiles_rec = None
    if world > 1 and not tiles_mode and not args.no_tiles:
        r.close()
        done = threading.Event()

        def bail():  # a hung collective must not cost the run its JSON line: give up on the sub-record
            if not done.wait(240.0):
                if rank == 0:
                    print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": Wm,
                                      "ms_per_step": total_ms_max / (K * blocks), "higher_is_better": True, "scaling": "weak",
                                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, scene),
                                      "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                                      "gpu_launches": int(launches), "tiles": {"error": "sub-record timed out after 240 s"}}), flush=True)
                os._exit(0)
        threading.Thread(target=bail, daemon=True).start()
        try:
            tiles_rec = tiles_subrecord(args, rank, world, local)
        except Exception as e:  # recorded, never fatal for the headline line
            tiles_rec = {"error": repr(e)[:300]}
        done.set()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_kind = measured_peaks()
    alg = {
        "geometry": 36 * scene.n_vertices + 12 * scene.n_triangles + 192,  # mesh + indices + matrices
        "tile": scene.texture.texels.nbytes + 4 * W * own_rows,  # texture + this rank's rows of the image
    }
    kt = {"geometry": stage_avg["geometry_ms"], "tile": stage_avg["tile_ms"]}
    dom = max(kt, key=kt.get)
    achieved = alg[dom] / (kt[dom] / 1e3) / 1e9
    traffic, ncu_detail, tj = None, None, None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            traffic = tj.get(dom)
            det = tj.get("_detail", {}).get("tile" if dom == "tile" else "geom", {})
            # the path is issue/latency bound, not HBM bound: quote the issue-slot utilisation ncu saw
            ncu_detail = {"issue_active_pct": det.get("issue_active_pct"), "warp_instructions": det.get("warp_inst"),
                          "registers": det.get("regs"), "source": tj.get("_source"), "measured_in_run": False,
                          "capture_commit": tj.get("_commit")}
        except Exception:
            traffic = None
    frame_alg = scene.algorithmic_bytes()
    ms_per_step = total_ms_max / (K * blocks)
    roofline = {
        "bound": "hbm", "kernel": f"{dom}_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic,
        "traffic_source": {"measured_in_run": False, "what": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel from one "
                           "ncu --set full capture of the same command (profiles/traffic.json); ncu flushes the caches before the "
                           "launch, so the bin entries (16 B), raster records (48 B) and shade records (32 B) the geometry stage has "
                           "just written are counted as DRAM reads here, while in a running frame they are L2 hits",
                           "capture_commit": (tj or {}).get("_commit")},
        "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
        "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": kt[dom],
        "kernel_ms_all": stage_avg, "kernel_share_of_frame": kt[dom] / max(stage_avg["total_ms"], 1e-9),
        "frame_algorithmic_bytes": frame_alg, "frame_frac": frame_alg / (ms_per_step / 1e3) / 1e9 / peak,
        "timing": "per-stage CUDA events on the launch stream over profiled frames run right after the timed region",
        "ncu": ncu_detail,
        "note": "no dense contraction on this path (no tensor cores); DRAM throughput is < 3 % in ncu, the kernels are "
                "bound by instruction issue and latency of the exact non-FMA f32 arithmetic (DESIGN.md section 4)",
    }
    # secondary bound: algorithmic f32 operations of the reference algorithm (FMA is off:
    # 1 flop/lane/clock) from the work counters, against 148 SMs x 128 lanes x the max SM clock
    cpf = {k: v / (K * blocks) for k, v in cnt.items()}
    f_alg = (28 * scene.n_vertices + 40 * cpf["n_tris_in"] + 60 * cpf["n_tris_setup"] + 72 * cpf["n_bbox_px"]
             + 30 * cpf["n_samples_written"] + 180 * cpf["n_shaded_px"])
    sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
    roofline["fp32"] = {"algorithmic_gflop_per_frame": f_alg / 1e9, "achieved": f_alg / (ms_per_step / 1e3) / 1e12,
                        "peak": fp32_peak, "unit": "TFLOP/s (non-FMA f32)", "frac": f_alg / (ms_per_step / 1e3) / 1e12 / fp32_peak,
                        "formula": "28*Nv + 40*Nt_in + 60*Nt_setup + 72*N_bbox_px + 30*N_samples + 180*N_shaded_px (texture FS)",
                        "note": "the reference's arithmetic, not the GPU's: 72 flops per bbox pixel are credited for EVERY triangle, "
                                "also the back-facing ones (62 % of this frame's bbox pixels) that the GPU path culls exactly without "
                                "walking them; the useful flops actually executed are about half of the numerator"}
    # third reading: the issue-slot roofline.  ncu counted the warp instructions one C2 frame executes (all kernels,
    # profiles/traffic.json); a B200 issues at most 148 SMs x 4 schedulers x clock of them per second.
    try:
        det = json.loads(tp.read_text()).get("_detail", {}) if (tp.exists() and not tiles_mode and args.n_phi == 1001 and args.n_theta == 501) else {}
        winst = sum(float(v.get("warp_inst", 0.0)) for v in det.values())
        if winst > 0:
            issue_peak = 148 * 4 * sm_mhz * 1e6
            roofline["issue"] = {"warp_instructions_per_frame": winst, "achieved_ginst_per_s": winst / (ms_per_step / 1e3) / 1e9,
                                 "peak_ginst_per_s": issue_peak / 1e9, "frac": winst / (ms_per_step / 1e3) / issue_peak,
                                 "measured_in_run": False,
                                 "source": "smsp__inst_executed.sum per kernel from profiles/traffic.json (ncu --set full, same workload)"}
    except Exception:
        pass
    cb, cb_c1, other_configs = None, None, []
    if not args.no_cpu_baseline and not tiles_mode and n_gpus == 1:  # rank 0 at N=1 only (bounded sample)
        cb = cpu_baseline(scene, budget_s=args.cpu_budget)
        # BASELINE configs[0]: the crate's default scene (main.rs:93-105, 1280x720, cube + sphere, 268 triangles) -- the config
        # BASELINE names "on CPU" -- next to the same frame on the GPU
        c1 = scenes.default_scene(1.0)
        cb_c1 = cpu_baseline(c1, budget_s=3.0)
        r1 = Renderer(c1.width, c1.height, device=local)
        r1.uniforms().bind_texture(0, c1.texture)
        dms = [r1.upload(d.mesh) for d in c1.draws]
        ts = []
        for i in range(12):
            scenes.render_scene(r1, c1, dms)
            r1.framebuffer_device()
            ts.append(r1.timings()["total_ms"])
        r1.close()
        cb_c1["gpu_ms_per_frame"] = statistics.median(ts[2:])
        cb_c1["gpu_mtris_per_s"] = c1.n_triangles / cb_c1["gpu_ms_per_frame"] / 1e3
        cb_c1["workload"] = "BASELINE configs[0]: default `cargo run --release` scene at elapsed = 1.0 (main.rs:93-105), 1280x720, 4xMSAA"
        # the other single-GPU BASELINE configs, device-resident meshes, per-stage CUDA events (median of 5 frames): parity
        # of each at FULL size is tests/test_gpu_fullsize.py; they are reported here so the driver's record carries them
        for cname, mk in (("configs[2] C3: 250K triangles straddling the near plane, 3840x2160", scenes.near_clip_scene),
                          ("configs[3] C4(i): 1M-triangle sphere on 8192x8192, one GPU", lambda: scenes.sphere_scene(width=8192, height=8192)),
                          ("configs[3] C4(ii): clipped full-screen quad on 8192x8192, one GPU", lambda: scenes.fullscreen_quad_scene(8192, 8192))):
            sc = mk()
            rc = Renderer(sc.width, sc.height, device=local)
            rc.uniforms().bind_texture(0, sc.texture)
            dms = [rc.upload(d.mesh) for d in sc.draws]
            ts = []
            for i in range(8):
                if i == 3:
                    rc.reset_counters()
                scenes.render_scene(rc, sc, dms)
                rc.framebuffer_device()
                ts.append(rc.timings())
            t = {k: statistics.median(x[k] for x in ts[3:]) for k in ts[0]}
            cc = rc.counters()
            rc.close()
            other_configs.append({"config": cname, "triangles": sc.n_triangles, "ms": t, "mtris_per_s": sc.n_triangles / t["total_ms"] / 1e3,
                                  "gsamples_per_s": cc["n_samples_written"] / 5 / t["total_ms"] / 1e6,
                                  "hbm_floor_frac": sc.algorithmic_bytes() / (t["total_ms"] / 1e3) / 1e9 / peak})

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": Wm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if tiles_mode else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, scene),
        "gsamples_per_s": gsamples, "ms_per_frame": ms_per_step, "frame_latency_ms": frame_latency_ms,
        "gather": (args.gather if tiles_mode and n_gpus > 1 else None),
        "interleave_band_tile_rows": (args.band if tiles_mode and n_gpus > 1 and args.gather == "peer" else None), "assembled_frame_matches_single_gpu": assembled_ok,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(e2.item()) / e2e_steps * 1e3, "steps": e2e_steps,
                "sync_call_latency_ms": e2e_sync_ms,
                "copies_only_ms_per_step": copy_ms, "pcie_ceiling_frac": copy_ms / (float(e2.item()) / e2e_steps * 1e3),
                "resident_mesh": e2e_resident,
                "copies_only_gb_per_s_all_ranks": (h2d + d2h) * n_gpus / (copy_ms / 1e3) / 1e9,
                "host_binding": host_binding,
                "how": "rz_render_host (pinned host mesh, H2D on the upload stream) + rz_framebuffer_host_async (D2H of the "
                       "image on the download stream) every step; copies of neighbouring steps overlap the kernels, "
                       "wall clock over all steps incl. the final rz_sync"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "tiles": tiles_rec,
        "cpu_baseline": cb,
        "cpu_baseline_c1": cb_c1,
        "other_configs": other_configs or None,
        "counters_per_frame": cpf,
        "timed": {"blocks_of_K_steps": blocks, "frames": K * blocks, "device_ms": total_ms_max,
                  "block_ms_min_median_max": ([min(block_ms), statistics.median(block_ms), max(block_ms)] if block_ms else None),
                  "how": "the K-step block repeated back to back until >= %.2f s are timed; one CUDA-event pair around all "
                         "blocks (barrier + synchronize on both sides), max over ranks; block spread from an event after "
                         "every block (frames of neighbouring blocks overlap, so single blocks are approximate)" % args.min_timed_s},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
