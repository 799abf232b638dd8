#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame of the per-tile tracing hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1: one rank per GPU)

A step is one frame of the workload (default: killeroo + ground plane, 3840x2160, 16 spp, the
north-star target scene; --workload C1..C5 selects the BASELINE.json configs).  At N > 1 the frame
is tile-sharded: every rank holds the replicated scene + grid, renders the strips
strip_id % N == rank and stores its pixels straight into rank 0's device framebuffer through a
CUDA-IPC mapping (NVLink peer stores); no collective is on the data path.

Prints ONE JSON line (rank 0).  `value` = primary rays of the whole frame / device time of the
trace kernel (CUDA events on its stream, max over ranks per step, L2 flushed before every step);
`e2e` = the same through the reference-facing call cuda_trace_tiles() with a HOST framebuffer
(tile list H2D + framebuffer D2H inside the timed region).  `--impl reference` times the
reference's own threaded CPU renderer (oracle/_ref) on the host cores instead.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "cpp-11-ray-trace-march-framework_b200"

METRIC = "Mrays/s"


def pkg(sub):
    return importlib.import_module(PKG + "." + sub)


def workload(name):
    scenes = pkg("scenes")
    if name not in scenes.CONFIGS:
        raise SystemExit("unknown workload %r (have: %s)" % (name, ", ".join(scenes.CONFIGS)))
    scene, w, h, spp, res = scenes.CONFIGS[name]
    return dict(name=name, scene=scene, width=w, height=h, spp=spp, grid_res=res)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p.get("hbm_gbs", 6650.0)), "measured", float(p.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


# ---------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- reference (CPU)
def time_reference_cpu(wl, steps, warmup, budget_s):
    """The reference's own CPU renderer (oracle/_ref: unmodified sources + headless driver) on all
    host threads.  A step renders every `stride`-th tile of the reference's 12x9 layout through
    Renderer::RenderTile with hardware_concurrency() threads (stride 1 = the reference's own
    WorkerThread pool on the whole frame); stride is chosen so the run fits budget_s.
    Falls back to the oracle port only if oracle/_ref could not be built."""
    from oracle import pyoracle as po
    scenes = pkg("scenes")
    w, h, spp = wl["width"], wl["height"], wl["spp"]
    # The 50 M-triangle soup cannot go through the reference's own grid build in bounded time: its
    # FLT_MIN-seeded candidate ranges (triangle.h:123) make it test ~10^5 cells per triangle at 512^3.
    # For that scene the CPU side is the oracle port with tight candidate ranges (identical grid).
    use_port = not po.have_ref() or wl["scene"].startswith("tiger_soup")
    if not use_port:
        ref = po.Ref.get()
        threads = ref.hardware_threads()
        mesh, fov, cam = scenes.build(ref.api, wl["scene"])
        t0 = time.perf_counter()
        r = ref.renderer(mesh, fov, cam, wl["grid_res"])
        build_s = time.perf_counter() - t0
        # probe: 1/9 of the tiles
        sec, tiles, pix = r.render_tile_subset(w, h, spp, 9, 0, threads)
        est_full = sec * (w * h) / max(pix, 1)
        per_step_budget = budget_s / max(steps + warmup, 1)
        stride = 1
        while stride < 9 and est_full / stride > per_step_budget:
            stride += 1
        times, rays = [], 0
        for i in range(warmup + steps):
            if stride == 1:
                sec, _ = r.render(w, h, spp, want_image=False)
                pix = w * h
            else:
                sec, tiles, pix = r.render_tile_subset(w, h, spp, stride, i % stride, threads)
            if i >= warmup:
                times.append(sec)
                rays += pix * spp
        total = sum(times)
        sample = ("%d of the 108 tiles per step (every %d-th tile, offset rotating), %d steps" % (108 // stride, stride, steps)
                  if stride > 1 else "whole frame through the reference's own worker pool, %d steps" % steps)
        return dict(value=rays / total / 1e6, unit=METRIC, cores=threads, kind="reference", sample=sample,
                    ms_per_step=1e3 * total / len(times), ms_per_frame_est=1e3 * total / rays * (w * h * spp),
                    grid_build_s=build_s)
    # port fallback
    port = po.Port.get()
    host = pkg("hostapi").host_api()
    mesh, fov, cam = scenes.build(host, wl["scene"])
    vtx, tri = mesh.arrays()
    threads = os.cpu_count() or 1
    ps = port.scene(vtx, tri, wl["grid_res"], n_threads=threads, tight_ranges=True)
    rows = max(8, h // 64) if len(tri) > 1000000 else max(8, h // 16)
    times, rays = [], 0
    for i in range(warmup + steps):
        y0 = (i * rows) % max(h - rows, 1)
        t0 = time.perf_counter()
        ps.render(cam, fov, w, h, spp, y_begin=y0, y_end=y0 + rows, n_threads=threads)
        sec = time.perf_counter() - t0
        if i >= warmup:
            times.append(sec)
            rays += rows * w * spp
    total = sum(times)
    return dict(value=rays / total / 1e6, unit=METRIC, cores=threads, kind="port",
                sample="%d image rows per step, %d steps (oracle port: %s)" % (
                    rows, steps, "reference grid build unbounded for this scene" if po.have_ref() else "oracle/_ref not built"),
                ms_per_step=1e3 * total / len(times), ms_per_frame_est=1e3 * total / rays * (w * h * spp))


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    base = time_reference_cpu(wl, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
        "ms_per_frame_est": base["ms_per_frame_est"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic camera over the reference's own mesh assets",
        "config": {"workload": wl["name"], "scene": wl["scene"], "width": wl["width"], "height": wl["height"],
                   "spp": wl["spp"], "grid_res": wl["grid_res"]},
        "cpu_baseline": {"value": base["value"], "unit": METRIC, "cores": base["cores"], "kind": base["kind"],
                         "sample": base["sample"]},
        "e2e": {"value": base["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU (ours)
def run_ours(args, wl, rank, world, local_rank, dist):
    capi = pkg("capi")
    hostapi = pkg("hostapi")
    scenes = pkg("scenes")
    w, h, spp = wl["width"], wl["height"], wl["spp"]
    rays_per_frame = w * h * spp

    # scene through the product's own host library (Mesh / Matrix44f mirror), then onto the GPU
    host = hostapi.host_api()
    mesh, fov, cam = scenes.build(host, wl["scene"])
    vtx, tri = mesh.arrays()
    ct = capi.CudaTrace(devices=[local_rank])
    t0 = time.perf_counter()
    ct.upload_scene(vtx, tri, wl["grid_res"])
    upload_s = time.perf_counter() - t0
    ct.set_shard(rank, world)
    fov_xs, aspect = host.camera_constants(fov, w, h)
    frame = ct.make_frame(w, h, spp, cam, fov_xs, aspect)
    rects = capi.CudaTrace.make_tiles(capi.full_frame_tiles(w, h))  # converted once: the layout repeats every frame

    # rank 0 owns the framebuffer; the others map it (CUDA IPC) and store into it over NVLink
    multirank = pkg("multirank")
    group = multirank.RankGroup(dist, "cuda" if dist is not None else None)
    multirank.share_framebuffer(ct, group, w, h)
    barrier = group.barrier

    pinned = capi.PinnedImage(w, h)  # page-locked host framebuffer: the D2H copy is one DMA
    host_fb = pinned.array

    # warm-up (also first-touch of sample table, framebuffer, host registration)
    for _ in range(max(args.warmup, 3)):
        ct.trace_tiles_async(frame, rects)
        ct.sync()
    barrier()
    if rank == 0:
        ct.read_framebuffer(host_fb)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-timed region: K frames, L2 flushed before each, CUDA events around the kernel
    launches0 = ct.kernel_launches()
    barrier()
    step_ms = np.zeros(args.steps, np.float64)
    for i in range(args.steps):
        ct.flush_l2()
        ct.trace_tiles_async(frame, rects)
        ct.sync()
        step_ms[i] = ct.last_kernel_ms()
    barrier()
    launches = ct.kernel_launches() - launches0
    step_ms = group.allreduce_max(step_ms)      # a frame is done when its slowest rank is
    launches = int(group.allreduce_sum([launches])[0])
    total_ms = float(step_ms.sum())

    # ---- end-to-end region: the reference-facing call with a host framebuffer
    if world > 1:
        # every rank now also counts its finished strips per row band behind rank 0's framebuffer
        # (system-scope release, a few % of kernel time), so that rank 0 can ship bands to the host
        # while the frame is still being traced and needs no barrier before the read-back
        ct.set_shard_signals(True)
        barrier()
    e2e_s = 0.0
    e2e_kernel_ms = np.zeros(args.steps, np.float64)   # the trace kernel inside the end-to-end region (max over ranks)
    e2e_marks = np.zeros(7, np.float64)                # rank 0: submitted / traced / copied / returned, ms since call entry
    for i in range(args.steps):
        ct.flush_l2()
        ct.sync()
        barrier()
        t0 = time.perf_counter()
        if rank == 0:
            # the reference-facing call: returns when the whole frame is in the HOST buffer.  Row bands
            # are copied out as soon as every rank's strips of that band are finished (no barrier
            # between tracing and read-back)
            ct.trace_tiles(frame, rects, out=host_fb)
        else:
            ct.trace_tiles_async(frame, rects)
            ct.sync()
        barrier()
        e2e_s += time.perf_counter() - t0
        e2e_kernel_ms[i] = ct.last_kernel_ms()
        if rank == 0:
            e2e_marks += np.array(ct.last_call_timing()) / args.steps
    e2e_s = float(group.allreduce_max([e2e_s])[0])
    mine = np.zeros(world, np.float64)
    mine[rank] = e2e_kernel_ms.mean()
    e2e_kernel_by_rank = group.allreduce_sum(mine)     # diagnostics: each rank's kernel inside the end-to-end region
    e2e_kernel_ms = group.allreduce_max(e2e_kernel_ms)
    t0 = time.perf_counter()
    for _ in range(20):
        barrier()
    barrier_ms = float(group.allreduce_max([(time.perf_counter() - t0) / 20 * 1e3])[0])
    clocks = sampler.stop() if rank == 0 else None

    # ---- work counters of this frame (untimed, instrumented kernel) for the roofline figures
    counters = None
    if world == 1:
        ct.set_counting(True)
        ct.trace_tiles_async(frame, rects)
        ct.sync()
        counters = ct.get_counters()
        ct.set_counting(False)

    if rank != 0:
        ct.close()
        return

    value = rays_per_frame * args.steps / (total_ms * 1e-3) / 1e6
    e2e_value = rays_per_frame * args.steps / e2e_s / 1e6
    hbm_peak, peak_kind, sm_max = measured_peaks()
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic camera over the reference's own mesh assets (assets/meshes)",
        "config": {"workload": wl["name"], "scene": wl["scene"], "width": w, "height": h, "spp": spp,
                   "grid_res": wl["grid_res"], "triangles": int(len(tri)), "rays_per_frame": rays_per_frame,
                   "tiles": "12x9 (reference layout); ~128-ray pixel-block strips, chunks of 32 strips dealt round-robin over ranks",
                   "l2": "flushed before every timed step (256 MiB memset, untimed); scene itself is L2-resident",
                   "scene_upload_and_grid_build_s": upload_s},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": METRIC, "ms_per_step": 1e3 * e2e_s / args.steps,
                "kernel_ms_per_step": float(e2e_kernel_ms.mean()),
                "kernel_ms_by_rank": [round(float(x), 4) for x in e2e_kernel_by_rank], "barrier_ms": barrier_ms,
                "rank0_call_ms": {"prepared": e2e_marks[4], "launching": e2e_marks[5], "launched": e2e_marks[6],
                                  "submitted": e2e_marks[0], "traced": e2e_marks[1], "copied": e2e_marks[2], "returned": e2e_marks[3]},
                "h2d_bytes_per_step": 108 * 16 + 109 * 4, "d2h_bytes_per_step": w * h * 4},
        "gpu_launches": int(launches),
        "step_ms": [round(float(x), 4) for x in step_ms],
    }
    if counters:
        # SURVEY.md section 8(d) (DESIGN.md "Algorithmic work"): per frame
        #   bytes = 8*C + 40*T + 48*Hh + 4*P      flops = 80*R + 4*C + 40*T + 36*Hh + 4*(R-Hh) + 12*P
        R, Cc, T, Hh, P = counters["rays"], counters["cells"], counters["tri_tests"], counters["hits"], w * h
        alg_bytes = 8 * Cc + 40 * T + 48 * Hh + 4 * P
        alg_flops = 80 * R + 4 * Cc + 40 * T + 36 * Hh + 4 * (R - Hh) + 12 * P
        sec = total_ms * 1e-3 / args.steps
        traffic = None
        prof = os.path.join(ROOT, "profiles", "latest_traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    traffic = json.load(f).get(wl["name"])
            except (OSError, ValueError):
                traffic = None
        fp32_peak = 148 * 128 * sm_max * 1e6 / 1e12  # non-FMA instr-flop/s at max clock (kernel runs -fmad=false)
        line["roofline"] = {
            "bound": "hbm", "achieved": alg_bytes / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_bytes / sec / 1e9 / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
            "algorithmic_bytes_per_launch": alg_bytes,
            "note": "algorithmic bytes are L1/L2 request traffic for this L2-resident scene, not DRAM traffic; "
                    "the kernel is FP32-issue/latency bound, see fp32",
            "fp32": {"achieved_tflops": alg_flops / sec / 1e12, "peak_tflops_nonfma": fp32_peak,
                     "frac": alg_flops / sec / 1e12 / fp32_peak, "algorithmic_flops_per_launch": alg_flops},
            "counters": counters,
        }
        try:
            # ceilings measured on this box right now (csrc/peaks.cu): FP32 without FMA, L2 read bandwidth
            fp32_meas, l2_meas = capi.measure_peaks(local_rank)
            line["roofline"]["fp32"].update({"peak_tflops_nonfma_measured": fp32_meas,
                                             "frac_of_measured": alg_flops / sec / 1e12 / fp32_meas})
            line["roofline"]["l2"] = {"achieved": alg_bytes / sec / 1e9, "peak_measured": l2_meas, "unit": "GB/s",
                                      "frac": alg_bytes / sec / 1e9 / l2_meas,
                                      "note": "algorithmic bytes against the measured L2 read bandwidth (SURVEY 8d: the "
                                              "bandwidth roofline of the L2-resident configs)"}
        except Exception as e:  # the figures above stand without it
            line["roofline"]["fp32"]["peak_measured_error"] = str(e)
    if world == 1 and not args.no_cpu_baseline:
        base = time_reference_cpu(wl, 1, 0, budget_s=25.0)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["ms_per_frame_est"] = base["ms_per_frame_est"]
    ct.close()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="killeroo4k")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.steps = max(args.steps, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.workload)

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU); see the module docstring" % args.gpus)
    try:
        run_ours(args, wl, rank, world, local_rank, dist)
    finally:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
