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

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period, self.rows, self.proc, self.thread, self.stop_flag = index, period_s, [], None, None, False
        self.mode = None

    def _reasons(self, mask):
        import pynvml as N

        bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
        return [n for n, b in bits.items() if mask & b]

    def _poll(self):
        import pynvml as N

        while not self.stop_flag:
            try:
                sm = N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)
                mask = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((time.time(), float(sm), float(self.max_sm), self._reasons(mask)))
            except Exception:
                pass
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
    if world > 1:
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
            bj.view = views[0]
            rj.render(mj[0], 0, 0)
            rj.framebuffer_device()  # sizes the device buffers
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
    t_wall0 = time.time()
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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(K):
            rj, sj, bj, mj = lanes[s % L]
            bj.view = views[Wm + s]
            rj.render(mj[(s // L) % len(mj)], 0, 0)
            rj.framebuffer_async()
        for rj, sj, bj, mj in lanes[1:]:
            done = torch.cuda.Event()
            done.record(sj)
            stream.wait_event(done)
        e1.record(stream)
        for x in lanes:
            x[0].sync()
    torch.cuda.synchronize()
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
    frames_total = K * (1 if tiles_mode else n_gpus)
    value = (scene.n_triangles * frames_total) / (total_ms_max / 1e3) / 1e6
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
    e2 = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    e2e_value = scene.n_triangles * e2e_steps * (1 if tiles_mode else n_gpus) / float(e2.item()) / 1e6
    h2d = mesh.vertices.nbytes + mesh.attributes.nbytes + mesh.indices.nbytes + 192
    d2h = W * H * 4

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
    traffic, ncu_detail = None, None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            traffic = tj.get(dom)
            det = tj.get("_detail", {}).get("tile" if dom == "tile" else "geom", {})
            # the path is issue/latency bound, not HBM bound: quote the issue-slot utilisation ncu saw
            ncu_detail = {"issue_active_pct": det.get("issue_active_pct"), "warp_instructions": det.get("warp_inst"),
                          "registers": det.get("regs"), "source": tj.get("_source")}
        except Exception:
            traffic = None
    frame_alg = scene.algorithmic_bytes()
    ms_per_step = total_ms_max / K
    roofline = {
        "bound": "hbm", "kernel": f"{dom}_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
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
    cpf = {k: v / K for k, v in cnt.items()}
    f_alg = (28 * scene.n_vertices + 40 * cpf["n_tris_in"] + 60 * cpf["n_tris_setup"] + 72 * cpf["n_bbox_px"]
             + 30 * cpf["n_samples_written"] + 180 * cpf["n_shaded_px"])
    sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
    roofline["fp32"] = {"algorithmic_gflop_per_frame": f_alg / 1e9, "achieved": f_alg / (ms_per_step / 1e3) / 1e12,
                        "peak": fp32_peak, "unit": "TFLOP/s (non-FMA f32)", "frac": f_alg / (ms_per_step / 1e3) / 1e12 / fp32_peak,
                        "formula": "28*Nv + 40*Nt_in + 60*Nt_setup + 72*N_bbox_px + 30*N_samples + 180*N_shaded_px (texture FS)"}
    # third reading: the issue-slot roofline.  ncu counted the warp instructions one C2 frame executes (all kernels,
    # profiles/traffic.json); a B200 issues at most 148 SMs x 4 schedulers x clock of them per second.
    try:
        det = json.loads(tp.read_text()).get("_detail", {}) if (tp.exists() and not tiles_mode and args.n_phi == 1001 and args.n_theta == 501) else {}
        winst = sum(float(v.get("warp_inst", 0.0)) for v in det.values())
        if winst > 0:
            issue_peak = 148 * 4 * sm_mhz * 1e6
            roofline["issue"] = {"warp_instructions_per_frame": winst, "achieved_ginst_per_s": winst / (ms_per_step / 1e3) / 1e9,
                                 "peak_ginst_per_s": issue_peak / 1e9, "frac": winst / (ms_per_step / 1e3) / issue_peak,
                                 "source": "smsp__inst_executed.sum per kernel from profiles/traffic.json (ncu --set full, same workload)"}
    except Exception:
        pass
    cb = None
    if not args.no_cpu_baseline and not tiles_mode and n_gpus == 1:  # rank 0 at N=1 only (bounded sample)
        cb = cpu_baseline(scene, budget_s=args.cpu_budget)

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
                "how": "rz_render_host (pinned host mesh, H2D on the upload stream) + rz_framebuffer_host_async (D2H of the "
                       "image on the download stream) every step; copies of neighbouring steps overlap the kernels, "
                       "wall clock over all steps incl. the final rz_sync"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cb,
        "counters_per_frame": {k: v / K for k, v in cnt.items()},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
