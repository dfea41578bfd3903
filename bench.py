#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: ms/frame for 6 M Gaussians, 1920x1080, SH degree 3.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one frame of the hot path (preprocess+scan -> depth sort of the visible Gaussians -> duplication -> tile sort of
the pairs -> ranges -> blend). At N > 1
(launched under torchrun, one rank per GPU) the Gaussian set is replicated with one NCCL broadcast, independent camera
views are sharded round-robin across ranks (view = step * N + rank) and every finished frame is pushed, inside the timed
region, into its slot of one frame array in rank 0's HBM (CUDA IPC mapping; a copy engine moves it over NVLink while the next
frame renders; one stream-ordered fence per K steps) — weak scaling; `value` = max-over-ranks time / total frames.

JSON keys beyond the base contract: `roofline` (the onesweep pass kernel of the tile sort against measured HBM peak), `cpu_baseline`
(the CPU oracle on the host cores), `e2e` (camera on the host -> rasterFrame -> draw() into pinned host memory),
`stages_ms`, `sort_gkeys_per_s`, `pairs`, `visible`, `clocks`, `rounds` / `timed_region_s` (the K-step loop is repeated until the
timed region holds >= 1 s of device time; `value` is the mean over all of its frames), `config5` (BASELINE configs[4]: the 64-view
batch of a 3 M-Gaussian scene, strong scaling over the N GPUs, views/s).

`--impl reference`: the reference's Slang path cannot run here (no slangc, no Vulkan ICD); the reference
arm is the CPU restatement of those shaders (oracle/, kind "port") on all host cores, same workload, same metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SH_DEGREE = 1920, 1080, 3
N_GAUSSIANS = 6_000_000
SCENE_SEED = 3
LOG_SCALE_MEAN = -5.2          # calibrated: P/N = 2.64 at 1080p (reference bicycle scene: 15.7 M pairs, demo/README.md:20)
RING_VIEWS = 64                # cameras on a ring through the "garden" eye (2.8, 2.8, 2.6)
RING_PHI, RING_RADIUS, RING_THETA0 = 0.9898, 4.7371, 0.7853982
MIN_TIMED_S = 1.05              # device time the timed region must hold: the K-step loop is repeated until it does
METRIC = "ms/frame 6M-Gaussian 1080p SH3"
WORKLOAD = "synthetic 6M Gaussians (MipNeRF360-garden scale) SH3 at 1920x1080"


def scene_cached(n):
    from torpedo_b200 import scenes
    path = f"/tmp/tpd_garden_{n}_{SCENE_SEED}_{LOG_SCALE_MEAN}.npy"
    if os.path.exists(path):
        try:
            g = np.load(path)
            if g.shape == (n, 60):
                return g
        except Exception:
            pass
    g = scenes.garden(n, SCENE_SEED, log_scale_mean=LOG_SCALE_MEAN)
    try:
        np.save(path, g)
    except Exception:
        pass
    return g


def ring_camera_params(view):
    theta = RING_THETA0 + 2.0 * np.pi * (view % RING_VIEWS) / RING_VIEWS
    return float(np.float32(theta)), RING_PHI, RING_RADIUS


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, name in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per onesweep launch and where the figure comes from: the committed summary of one `ncu --set full` capture
    (profiles/onesweep_traffic.json). bench.py cannot run under a profiler, so the number is NOT measured in this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "onesweep_traffic.json")) as f:
            d = json.load(f)
            return float(d["dram_bytes_per_launch"]), d.get("source", "profiles/onesweep_traffic.json")
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle on all host cores
# ---------------------------------------------------------------------------------------------------

def cpu_frames(g, ubo, frames, warmup):
    from oracle import oracle as O
    ft = O.FrameTimer(g, WIDTH, HEIGHT, SH_DEGREE, capacity=int(3.2 * len(g)) + 4096)
    for _ in range(warmup):
        ft.frame(ubo)
    t0 = time.perf_counter()
    stages, pairs = [], 0
    for _ in range(frames):
        pairs, ms = ft.frame(ubo)
        stages.append(ms)
    total = (time.perf_counter() - t0) * 1e3
    return total / max(frames, 1), pairs, stages[-1] if stages else {}, O.num_threads()


def garden_ubo():
    """Camera block of ring view 0 for the reference arm, built WITHOUT the product: the reference's own Camera.cpp /
    PerspectiveCamera.cpp (oracle/_ref, compiled in place) when that library travelled with the tree, else the same
    look-at in numpy float64 (differs in the last ulp at most: P moves by < 0.1 %)."""
    from oracle import oracle as O
    theta, phi, radius = ring_camera_params(0)
    if O.ref_available():
        eye = O.ref_to_cartesian(theta, phi, radius)
        return O.ref_camera_ubo(WIDTH, HEIGHT, [float(x) for x in eye], (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)), "oracle/_ref (reference Camera.cpp)"
    eye = np.array([radius * np.sin(phi) * np.cos(theta), radius * np.sin(phi) * np.sin(theta), radius * np.cos(phi)])
    fwd = -eye / np.linalg.norm(eye)                       # rendering/src/Camera.cpp:3-16: z forward, x right, y down
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0])); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = right, down, fwd
    view[:3, 3] = [-right @ eye, -down @ eye, -fwd @ eye]
    fy, near, far = np.sqrt(3.0), 0.01, 100.0              # extension/src/PerspectiveCamera.cpp:9-23, Camera.h:33-34
    proj = np.array([[fy * HEIGHT / WIDTH, 0, 0, 0], [0, fy, 0, 0], [0, 0, near / (near - far), near * far / (far - near)], [0, 0, 1, 0]])
    ubo = np.concatenate([view.ravel(), (proj @ view).ravel(), [proj[0, 0], proj[1, 1]]]).astype(np.float32)
    return ubo, "numpy look-at (oracle/_ref absent)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    O.use_all_cores()   # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm uses every core it may run on
    g = scene_cached(N_GAUSSIANS)
    ubo, camera_src = garden_ubo()
    # time-box: full 6 M frames cost ~2 s each on 8 cores; fall back to a prefix sample if K is very large
    n = N_GAUSSIANS
    scale = 1.0
    if (args.steps + args.warmup) * 2.5 > 420:
        n, scale = N_GAUSSIANS // 4, 4.0
    ms, pairs, stages, cores = cpu_frames(g[:n], ubo, args.steps, args.warmup)
    sample = (f"{args.steps} full frames of the {N_GAUSSIANS}-Gaussian workload" if scale == 1.0 else
              f"{args.steps} frames of the first {n} Gaussians, value scaled x{scale:g}")
    value = ms * scale
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ms/frame", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "n_gaussians": N_GAUSSIANS, "width": WIDTH, "height": HEIGHT, "sh_degree": SH_DEGREE,
                                        "impl_note": "CPU restatement of the reference's Slang shaders (oracle/), OpenMP on all host cores; lavapipe/slangc absent; camera: " + camera_src},
        "cpu_baseline": {"value": value, "unit": "ms/frame", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ms/frame", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pairs": pairs, "stages_ms": stages, "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------

def run_gpu(args):
    import torch
    import torch.distributed as dist

    from torpedo_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the rasterizer (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- scene: generated on rank 0, replicated with ONE NCCL broadcast -----------------------------
    eng = E.GaussianEngine(WIDTH, HEIGHT, device=local_rank)
    n = N_GAUSSIANS
    if world == 1:
        g = scene_cached(n)
        scene = E.Scene()
        scene.add_group(g)
        eng.compile(scene, E.Settings(SH_DEGREE))
    else:
        recs = torch.empty((n, 60), dtype=torch.float32, device=dev)
        if rank == 0:
            g = scene_cached(n)
            recs.copy_(torch.from_numpy(g))
        dist.broadcast(recs, src=0)
        torch.cuda.synchronize()
        eng.compile_device(recs.data_ptr(), n, E.Settings(SH_DEGREE), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        del recs
    stream = torch.cuda.current_stream().cuda_stream

    def ubo_for(view):
        cam = E.PerspectiveCamera(WIDTH, HEIGHT)
        cam.look_at(E.to_cartesian(*ring_camera_params(view)), (0, 0, 0), (0, 0, 1))
        return cam.pack()

    K, W = args.steps, max(args.warmup, 3)
    frame_bytes = WIDTH * HEIGHT * 4
    frames = torch.zeros((K, HEIGHT, WIDTH, 4), dtype=torch.uint8, device=dev)
    from torpedo_b200._lib import check, tpdcu
    lib = tpdcu()
    # N > 1: the K * N frames of a round are collected on rank 0, slot = step * N + rank of ONE array in rank 0's HBM that every
    # rank has mapped (torpedo_b200.multiview.SharedFrames; the mapping is made here, once, like the scene broadcast)
    shared, gathered, collect = None, None, "none (one GPU)"
    if world > 1:
        from torpedo_b200 import multiview as mv
        try:
            shared = mv.SharedFrames(K * world, HEIGHT, WIDTH, local_rank)
            collect = "copy-engine pushes over NVLink into one CUDA-IPC frame array on rank 0"
        except mv.SharedFramesUnavailable as e:   # raised on every rank alike: NCCL gathers instead, and the line says so
            gathered = [torch.zeros_like(frames) for _ in range(world)] if rank == 0 else None
            collect = f"NCCL gathers, 4 frames per asynchronous chunk (CUDA IPC frame array unavailable: {e})"
    GATHER_CHUNK = 4

    def render_steps(first_step, count, gather=False):
        pending = []
        for s in range(count):
            view = (first_step + s) * world + rank
            if shared is None:
                check(lib.tpdcu_bind_output_device_ptr(eng.ctx, frames[s].data_ptr(), WIDTH * 4))
                eng.raster_ubo(ubos[view % RING_VIEWS], SH_DEGREE, stream)
                if gather and world > 1 and ((s + 1) % GATHER_CHUNK == 0 or s + 1 == count):
                    c0 = s + 1 - ((s % GATHER_CHUNK) + 1)
                    pending.append(dist.gather(frames[c0:s + 1], [gt[c0:s + 1] for gt in gathered] if rank == 0 else None, dst=0, async_op=True))
            else:   # rendered into the engine's own target, pushed by a copy engine while the next frame renders
                eng.raster_ubo(ubos[view % RING_VIEWS], SH_DEGREE, stream)
                check(lib.tpdcu_read_frame_async(eng.ctx, shared.ptr_of_view(s * world + rank), WIDTH * 4, stream))
        for work in pending:
            work.wait()     # the current stream waits for the transfers; the host does not
        if shared is not None and gather:
            shared.fence()  # stream-ordered: behind it rank 0 holds every rank's frames of these steps

    ubos = [ubo_for(v) for v in range(RING_VIEWS)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also grows the pair buffers to their steady-state capacity: EVERY view of the ring, so that no frame of
    # the timed region can overflow them and be rendered truncated; `frames_repeated_in_timed_region` proves it) ---------
    for v in range(RING_VIEWS):
        eng.raster_ubo(ubos[v], SH_DEGREE, stream)
        eng.finish()
    render_steps(0, min(W, K), gather=True)
    eng.finish()
    barrier()

    # ---- timed region: R back-to-back rounds of K frames per rank (+ their collection on rank 0 at N > 1) -------------------
    # K frames are ~20 ms: one noisy neighbour or two clock samples would decide the headline. The K-step loop is therefore
    # repeated (R rounds, every round a different stretch of the camera ring) until the region between the two events holds
    # >= MIN_TIMED_S of device time; `value` is the mean over all R*K frames and `rounds` says how many there were.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    render_steps(W, K, gather=True)
    ev1.record()
    barrier()
    est = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    rounds = max(1, min(2000, int(np.ceil(MIN_TIMED_S * 1e3 / max(float(est.item()), 1e-3)))))
    repeats_before_timed = eng.frames_repeated()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0.record()
    for r in range(rounds):
        render_steps(W + r * K, K, gather=True)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    pairs, visible = eng.counts()
    cap_ok = pairs <= eng.capacity()
    eng.finish()
    repeated_in_timed = eng.frames_repeated() - repeats_before_timed   # a frame that overflowed its pair buffers was truncated
    t = torch.tensor([elapsed_ms, float(repeated_in_timed)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, repeated_in_timed = float(t[0].item()), int(t[1].item())

    # ---- per-stage times (CUDA events around every stage, separate untimed frames) ----------------------------------
    check(lib.tpdcu_bind_output_device_ptr(eng.ctx, None, 0))
    eng.enable_stage_timing(True)
    stage_runs = []
    for s in range(5):
        eng.raster_ubo(ubos[(s * world + rank) % RING_VIEWS], SH_DEGREE, stream)
        stage_runs.append((eng.stage_times_ms(), eng.counts()[0]))
    eng.enable_stage_timing(False)
    stages = {k: statistics.median(r[0][k] for r in stage_runs) for k in stage_runs[0][0]}
    stage_pairs = statistics.median(r[1] for r in stage_runs)
    sort_info = eng.sort_info()

    # ---- end-to-end through the public API with host buffers (rank-local; max over ranks) -----------------------------
    # The user's loop of the reference (demo/HelloGaussian/main.cpp:49-58): camera->lookAt on the host, rasterFrame, draw into
    # host memory. Like the reference's swap-chain loop it keeps frames in flight: frame k's copy to pinned host memory is
    # enqueued behind it and consumed while frame k+1 is already being rendered (two pinned buffers).
    NBUF = 3  # pinned host buffers: frame s-2 is consumed while frames s-1 and s are in flight
    hosts = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8).pin_memory() for _ in range(NBUF)]
    hosts_np = [hbuf.numpy() for hbuf in hosts]
    copied = [torch.cuda.Event() for _ in range(NBUF)]
    cam = E.PerspectiveCamera(WIDTH, HEIGHT)
    # like `value`, over enough frames that the fill and drain of the pipeline (one frame's latency + one copy, ~1 ms) and a
    # noisy neighbour do not decide the number: the K-step loop repeated until it holds >= ~0.5 s (e2e.steps says how many)
    e2e_steps = K * max(1, rounds // 2)
    serial_steps = min(K, 64)
    for s in range(3):
        cam.look_at(E.to_cartesian(*ring_camera_params(s * world + rank)), (0, 0, 0), (0, 0, 1))
        eng.raster_frame(cam, stream)
        eng.draw(hosts_np[0])
    repeats_before = eng.frames_repeated()
    barrier()
    checksum = 0
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        cam.look_at(E.to_cartesian(*ring_camera_params((W + s) * world + rank)), (0, 0, 0), (0, 0, 1))   # host: 136-byte camera block
        eng.raster_frame(cam, stream)                  # H2D (kernel argument) + frame, asynchronous
        eng.draw_async(hosts_np[s % NBUF], stream)     # D2H of the RGBA8 frame, enqueued behind the frame
        copied[s % NBUF].record()
        if s >= NBUF - 1:
            done_s = s - (NBUF - 1)
            copied[done_s % NBUF].synchronize()        # that frame is on the host now
            checksum += int(hosts_np[done_s % NBUF][HEIGHT // 2, WIDTH // 2, 0])
    torch.cuda.synchronize()                           # the last frames
    e2e_ms = (time.perf_counter() - t0) * 1e3
    checksum += int(hosts_np[(e2e_steps - 1) % NBUF][..., :3].sum())
    # and the same loop strictly serial (draw() blocks before the next rasterFrame): the latency of one frame end to end
    t0 = time.perf_counter()
    for s in range(serial_steps):
        cam.look_at(E.to_cartesian(*ring_camera_params((W + s) * world + rank)), (0, 0, 0), (0, 0, 1))
        eng.raster_frame(cam, stream)
        eng.draw(hosts_np[0])
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3
    repeats = eng.frames_repeated() - repeats_before
    t = torch.tensor([e2e_ms, e2e_serial_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_serial_ms = float(t[0].item()), float(t[1].item())

    if rank == 0:
        total_frames = K * world * rounds
        ms_per_frame = elapsed_ms / total_frames
        peak, peak_src = measured_peak_hbm()
        traffic, traffic_src = ncu_traffic_per_launch()
        # dominant HBM-bound kernel of the sort: one onesweep pass over the pair words (8 B in + 8 B out per pair)
        passes = max(int(stages["tile_passes_run"]), 1)
        pass_ms = (stages["tile_sort"] - stages["tile_sort_hist_plan"]) / passes
        alg_bytes = stage_pairs * 16.0
        achieved = alg_bytes / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else 0.0
        sort_ms = stages["depth_sort"] + stages["tile_sort"]
        depth_passes = int(stages["depth_passes_run"])
        line = {
            "metric": METRIC, "value": ms_per_frame, "unit": "ms/frame", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / (K * rounds), "rounds": rounds, "timed_region_s": elapsed_ms * 1e-3,
            "frames_repeated_in_timed_region": repeated_in_timed, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_gaussians": n, "width": WIDTH, "height": HEIGHT, "sh_degree": SH_DEGREE,
                       "views": f"ring of {RING_VIEWS} cameras through the garden eye, view = step*N + rank",
                       "parallelism": "one GPU" if world == 1 else f"views sharded over {world} GPUs, scene replicated (one NCCL broadcast); frames collected on rank 0 inside the timed region by: {collect}",
                       "pipelining": "3 frames in flight (the reference keeps 2, SurfaceRenderer.h:66): the memory-bound front of the next frames overlaps the SM-bound blend of the current one; single_frame_latency_ms is one frame alone",
                       "l2_policy": "inputs larger than L2 (scene arrays 1.4 GB, pair buffers 0.4 GB vs 126 MB L2); a different view every step"},
            "pairs": int(stage_pairs), "visible": int(visible), "capacity_ok": bool(cap_ok),
            "frames_in_flight": 3, "single_frame_latency_ms": stages["frame"],
            "stages_ms": stages,
            "sort_gkeys_per_s": stage_pairs / (sort_ms * 1e-3) / 1e9 if sort_ms > 0 else None,
            "sort": dict(sort_info, design="two-level: visible Gaussians by depth, duplication in depth order, pairs by tile",
                         bytes_per_pair=8 + 16 * passes, bytes_per_visible_gaussian=8 + 16 * depth_passes),
            # share of one frame alone per stage, and what bounds it (ncu: profiles/r2_ncu_frame_kernels.txt). The blend is the largest
            # stage and is bound by SM issue slots and L1 gathers, not by HBM or tensor throughput; the roofline object below
            # is therefore about the largest HBM-bound kernel, the onesweep pass (6 launches, a quarter of the frame).
            "stage_shares": {k: round(stages[k] / stages["frame"], 3) for k in ("preprocess", "depth_sort", "duplicate", "tile_sort", "ranges", "blend")},
            "stage_bounds": {"preprocess": "issue 60 % / HBM 49 % (the bit-exact chain is unfused fp32)", "depth_sort": "latency (L2-resident, two tiles per resident CTA)", "duplicate": "latency / LSU",
                             "tile_sort": "L1 data pipe 65-68 % (shared-memory wavefronts of ranking and scatter), HBM 40 %; histogram: issue", "ranges": "hbm", "blend": "SM issue 61 % + L1 gathers (on-demand SH colour inside)"},
            "roofline": {"kernel": "onesweep_kernel<WORDS> (one 8-bit digit pass of the tile sort over the pair words; average over the passes of a frame)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": pass_ms,
                         "peak_source": peak_src},
            "e2e": {"value": e2e_ms / e2e_steps / world, "unit": "ms/frame", "h2d_bytes_per_step": 136, "d2h_bytes_per_step": frame_bytes,
                    "steps": e2e_steps, "checksum": checksum, "serial_latency_ms": e2e_serial_ms / serial_steps,
                    "frames_repeated": repeats,
                    "scope": "rank-local: every rank delivers the frames of its own views into pinned host memory of the node (max over ranks / total frames); the collection on rank 0 that `value` includes is not on this path",
                    "note": "lookAt on host -> rasterFrame -> drawAsync into pinned host memory, frames in flight as in the reference's loop"},
            # per frame: setup, preprocess, duplication, ranges, blend order, blend, 2 x histogram (+ plan) + one onesweep launch per
            # 8-bit digit of the widest possible key of each sort: 4 for the 32 depth bits (a pass whose digit is constant
            # still launches and exits), ceil(tile_bits / 8) for the tiles
            "gpu_launches": K * rounds * (8 + 4 + ((((WIDTH + 15) // 16 * ((HEIGHT + 15) // 16) - 1).bit_length() + 7) // 8)),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            CPU_FRAMES = 16   # ~9 s of CPU work on 16 cores: a bounded sample of the same workload
            ms, cpu_pairs, cpu_stages, cores = cpu_frames(g, ubos[0], CPU_FRAMES, 1)
            line["cpu_baseline"] = {"value": ms, "unit": "ms/frame", "cores": cores, "kind": "port",
                                    "sample": f"{CPU_FRAMES} full frames (after 1 warm-up) of the same 6M-Gaussian scene, view 0", "stages_ms": cpu_stages}
    if shared is not None:
        if rank == 0:   # every slot of the last round arrived
            got = shared.tensor()
            collected_ok = all(int(got[v, ::16, ::16, :3].amax().item()) > 0 for v in range(K * world))
            line["config"]["frames_collected_on_rank0"] = f"{K * world} per round, all with pixels: {collected_ok}"
            del got
        shared.close()
    eng.close()
    del frames, gathered
    torch.cuda.empty_cache()
    config5 = None if args.no_config5 else run_config5(E, torch, dist, world, rank, local_rank, dev, args.config5_collect)
    if rank == 0:
        line["config5"] = config5
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_config5(E, torch, dist, world, rank, local_rank, dev, mode="push"):
    """BASELINE.json configs[4]: a 64-view batch of a 3 M-Gaussian scene at 1080p, views sharded round-robin over the ranks,
    scene replicated with one NCCL broadcast. STRONG scaling: the batch is fixed, N grows. Three ways of collecting the frames on
    rank 0 (torpedo_b200.multiview), all measured at N = 8 in one run (profiles/r2_collect_modes.jsonl): "push" (default) —
    render locally, a copy engine pushes every finished frame into its slot of rank 0's frame array over NVLink (CUDA IPC
    mapping, set up once like the scene broadcast), one stream-ordered fence per batch: 16.5 k views/s; "gather" — one NCCL
    gather per view slot in asynchronous chunks, whose receive kernels take SMs from rank 0's own views: 15.5 k; "direct" — the
    blend kernel stores its pixels straight into rank 0's array: 13.7 k (seven GPUs' 32-byte stores into one stall the senders).
    Returns the sub-object on rank 0 (max-over-ranks device time)."""
    from torpedo_b200 import multiview as mv
    from torpedo_b200 import scenes
    n, views, radius, chunk = 3_000_000, 64, 5.0, 2
    g = scenes.garden(n, 5, log_scale_mean=LOG_SCALE_MEAN) if rank == 0 else None
    recs = mv.broadcast_scene(torch.from_numpy(g).to(dev) if rank == 0 else None, n, dev)
    eng = E.GaussianEngine(WIDTH, HEIGHT, device=local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    eng.compile_device(recs.data_ptr(), n, E.Settings(SH_DEGREE), stream)
    torch.cuda.synchronize()
    del recs
    ubos = []
    for k in range(views):
        cam = E.PerspectiveCamera(WIDTH, HEIGHT)
        cam.look_at(E.to_cartesian(float(np.float32(2.0 * np.pi * k / views)), 0.9, radius), (0, 0, 0), (0, 0, 1))
        ubos.append(cam.pack())
    ubos = np.stack(ubos)

    from torpedo_b200._lib import check, tpdcu
    lib = tpdcu()

    def render_to(view_ids, ptrs):
        # asynchronous: every view is enqueued behind the previous one (three frames in flight inside the engine) and whatever
        # follows (fence or gather) is ordered behind the frames on the stream; the host only waits at the end of the batch
        for v, p in zip(view_ids, ptrs):
            check(lib.tpdcu_bind_output_device_ptr(eng.ctx, p, WIDTH * 4))
            eng.raster_ubo(ubos[v], SH_DEGREE, stream)

    def push_to(view_ids, ptrs):
        # render into the engine's own targets (local HBM) and let a copy engine push each finished frame into its slot of
        # rank 0's array: whole 8.29 MB transfers over NVLink, no SM on either side, overlapped with the next frames
        check(lib.tpdcu_bind_output_device_ptr(eng.ctx, None, 0))
        for v, p in zip(view_ids, ptrs):
            eng.raster_ubo(ubos[v], SH_DEGREE, stream)
            check(lib.tpdcu_read_frame_async(eng.ctx, p, WIDTH * 4, stream))

    def render_batch(view_ids, out):
        render_to(view_ids, [out[k].data_ptr() for k in range(len(view_ids))])

    shared, fallback = None, ""
    if mode in ("direct", "push"):
        try:
            shared = mv.SharedFrames(views, HEIGHT, WIDTH, local_rank)
        except mv.SharedFramesUnavailable as e:   # raised on every rank alike
            mode, fallback = "gather", f" (asked for the CUDA-IPC frame array, unavailable: {e})"
    mid = torch.cuda.Event(enable_timing=True)

    def timed_render(fn):
        def run(view_ids, ptrs):
            fn(view_ids, ptrs)
            mid.record()      # this rank's own frames are done here; what follows is the fence
        return run

    def batch():
        if shared is not None:
            mv.render_views_direct(timed_render(render_to if mode == "direct" else push_to), shared)
            return shared.tensor() if rank == 0 else None
        return mv.render_views(render_batch, views, HEIGHT, WIDTH, dev, chunk=chunk)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    check(lib.tpdcu_bind_output_device_ptr(eng.ctx, None, 0))
    for v in range(rank, views, world):                  # warm-up: pair buffers grow to the largest view of this rank
        eng.raster_ubo(ubos[v], SH_DEGREE, stream)
        eng.finish()
    for _ in range(2):
        batch()
        eng.finish()
    repeats_before = eng.frames_repeated()
    times, frames = [], None
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        frames = batch()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        own = torch.tensor([e0.elapsed_time(mid) if shared is not None else 0.0], dtype=torch.float64, device=dev)
        per_rank = [torch.zeros_like(own) for _ in range(world)]
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_gather(per_rank, own)
        else:
            per_rank = [own]
        times.append(float(t.item()))
        rank_ms = [round(float(x.item()), 3) for x in per_rank]
    eng.finish()
    repeated = eng.frames_repeated() - repeats_before
    out = None
    if rank == 0:
        ms = statistics.median(times)
        nonblack = [int((frames[v, ::8, ::8, :3].amax() > 0).item()) for v in range(views)]
        how = ("every rank's blend kernel stores its pixels straight into rank 0's frame array over NVLink (CUDA IPC mapping made once, "
               "outside the timed region, like the scene broadcast); one stream-ordered NCCL fence per batch" if mode == "direct" else
               f"one NCCL gather per view slot straight into view order, {chunk} slots per asynchronous chunk" + fallback)
        out = {"workload": "64-view batch of 3M Gaussians at 1080p sharded by view across N B200, frames collected on rank 0", "n_gaussians": n,
               "views": views, "n_gpus": world, "scaling": "strong", "ms_per_batch": ms, "ms_per_view": ms / views, "views_per_s": views / ms * 1e3,
               "batches_timed": len(times), "collect": mode, "gather": how,
               "own_frames_ms_per_rank_last_batch": rank_ms if shared is not None else None,
               "frames_repeated": repeated, "views_with_pixels": sum(nonblack),
               "limiter_at_n8": "8 views per GPU are a 3.76 ms job at the one-GPU rate; the rest is the fill and drain of the three-frame pipeline and the fence"}
    frames = None
    if shared is not None:
        shared.close()
    eng.close()
    return out


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries print on stdout too (NCCL writes "NCCL version ..." there when NCCL_DEBUG is set): keep the original stdout
    # for the JSON line and send everything else, C-level writes included, to stderr.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the 64-view x 3M multi-view batch (BASELINE configs[4]) sub-object")
    ap.add_argument("--config5-collect", default="push", choices=["direct", "push", "gather"],
                    help="how the multi-view batch's frames reach rank 0: copy-engine pushes or blend stores over NVLink into rank 0's array, or NCCL gathers")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
