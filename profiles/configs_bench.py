"""Every BASELINE.json config on one GPU: per-stage CUDA-event times + FULL-SIZE parity against the CPU oracle
(sorted keys, values, ranges bit-exact; image within 1 LSB). Writes gpurun_out/configs_r1.json.
Usage: python profiles/configs_bench.py [--skip-parity]"""
import json
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)
from torpedo_b200 import engine as E  # noqa: E402
from torpedo_b200 import scenes  # noqa: E402

skip_parity = "--skip-parity" in sys.argv


def camera(w, h, eye, center=(0, 0, 0), up=(0, 0, 1)):
    cam = E.PerspectiveCamera(w, h)
    cam.look_at(eye, center, up)
    return cam


def run(name, g, w, h, deg, cam, model=None, reps=20):
    scene = E.Scene()
    ent = scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene, E.Settings(deg))
    if model is not None:
        eng.transform(ent, model)
    for _ in range(3):
        eng.raster_frame(cam)
        eng.finish()
    eng.enable_stage_timing(True)
    runs = []
    for _ in range(reps):
        eng.raster_frame(cam)
        runs.append(eng.stage_times_ms())
    eng.enable_stage_timing(False)
    st = {k: round(statistics.median(r[k] for r in runs), 4) for k in runs[0]}
    # throughput with frames in flight: K back-to-back frames, one host sync at the end
    import torch
    for _ in range(4):
        eng.raster_frame(cam)
    eng.finish()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(40):
        eng.raster_frame(cam)
    eng.finish()
    torch.cuda.synchronize()
    throughput_ms = (time.perf_counter() - t0) * 1e3 / 40
    pairs, visible = eng.counts()
    res = {"config": name, "n": int(g.shape[0]), "w": w, "h": h, "sh": deg, "pairs": pairs, "visible": visible, "ms_per_frame_3_in_flight": round(throughput_ms, 4), "stages_ms": st,
           "sort": eng.sort_info(), "sort_gkeys_per_s": round(pairs / ((st["depth_sort"] + st["tile_sort"]) * 1e6), 2) if pairs else None}
    if not skip_parity:
        img = eng.draw()
        keys, vals = eng.read_sorted()
        ranges = eng.read_ranges()
        t0 = time.perf_counter()
        ref = O.render(g, cam.pack(), w, h, deg, models=None if model is None else np.asarray(model, np.float32).reshape(1, 16))
        res["cpu_oracle_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        res["cpu_cores"] = O.num_threads()
        diff = np.abs(img.astype(np.int32) - ref.rgba.astype(np.int32))
        mse = float(np.mean((img[..., :3].astype(np.float64) - ref.rgba[..., :3]) ** 2)) / 255.0 ** 2
        res["parity"] = {"pairs": pairs == ref.pairs, "keys": bool((keys == ref.keys).all()), "vals": bool((vals == ref.vals).all()),
                         "ranges": bool((ranges == ref.ranges).all()), "max_abs_rgb_lsb": int(diff[..., :3].max()),
                         "pixels_off_by_one": int((diff[..., :3].max(axis=-1) > 0).sum()), "alpha_255": bool((img[..., 3] == 255).all()),
                         "psnr_db": 99.0 if mse == 0 else round(-10 * np.log10(mse), 2)}
        lens = ref.ranges[:, 1] - ref.ranges[:, 0]
        res["tile_list"] = {"mean": float(lens.mean()), "max": int(lens.max())}
    eng.close()
    print(json.dumps(res), flush=True)
    return res


out = []
hello_eye = E.to_cartesian(0.785, 0.9, 8.0)
out.append(run("1a HelloGaussian literal (8192+1, SH0)", scenes.hello_gaussian(8192, seed=1), 1280, 720, 0, camera(1280, 720, hello_eye)))
out.append(run("1b HelloGaussian at BASELINE.json size (100k+1, SH3)", scenes.hello_gaussian(100000, seed=1), 1280, 720, 3, camera(1280, 720, hello_eye)))
g6 = bench.scene_cached(6_000_000)
out.append(run("2 synthetic 1M SH3 1080p", scenes.garden(1_000_000, 2, log_scale_mean=bench.LOG_SCALE_MEAN), 1920, 1080, 3, camera(1920, 1080, (2.8, 2.8, 2.6))))
out.append(run("3a synthetic 6M SH3 1080p (headline)", g6, 1920, 1080, 3, camera(1920, 1080, (2.8, 2.8, 2.6))))
out.append(run("3b synthetic 6M SH3 2160p", g6, 3840, 2160, 3, camera(3840, 2160, (2.8, 2.8, 2.6))))
out.append(run("4 VolumeSplatting dense 2M SH2 720p", scenes.dense_volume(2_000_000, seed=4), 1280, 720, 2,
               camera(1280, 720, (-2.0, -1.0, 0.0), (0, 0, 0), (0, -1, 0)), model=scenes.VOLUME_TRANSFORM))


def run_view_batch(name, g, w, h, deg, n_views, radius, check_views=(0, 21, 42, 63)):
    """BASELINE config 5 on one GPU: a batch of views of one scene through tpdcu_raster_views (frames stay in HBM), the
    batch timed with CUDA events; the sampled views are compared with the oracle image."""
    import torch
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene, E.Settings(deg))
    ubos = np.stack([camera(w, h, E.to_cartesian(2.0 * np.pi * k / n_views, 0.9, radius)).pack() for k in range(n_views)])
    frames = torch.zeros((n_views, h, w, 4), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):  # warm-up: buffers grow to the largest view
        eng.raster_views(ubos, frames.data_ptr(), h * w * 4, deg, stream)
        eng.finish()
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        eng.raster_views(ubos, frames.data_ptr(), h * w * 4, deg, stream)
        e1.record()
        eng.finish()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    res = {"config": name, "n": int(g.shape[0]), "w": w, "h": h, "sh": deg, "views": n_views, "ms_per_batch": round(ms, 3),
           "ms_per_view": round(ms / n_views, 4), "views_per_s": round(n_views / ms * 1e3, 1)}
    if not skip_parity:
        out_frames = frames.cpu().numpy()
        worst, off, pairs = 0, 0, []
        for k in check_views:
            ref = O.render(g, ubos[k], w, h, deg)
            d = np.abs(out_frames[k].astype(np.int32) - ref.rgba.astype(np.int32))
            worst = max(worst, int(d[..., :3].max()))
            off += int((d[..., :3].max(axis=-1) > 0).sum())
            pairs.append(int(ref.pairs))
        res["parity"] = {"views_checked": list(check_views), "max_abs_rgb_lsb": worst, "pixels_off_by_one": off, "pairs": pairs}
    eng.close()
    print(json.dumps(res), flush=True)
    return res


out.append(run_view_batch("5 64 views x 3M SH3 1080p, one GPU (ring, radius 5)", scenes.garden(3_000_000, 5, log_scale_mean=bench.LOG_SCALE_MEAN),
                          1920, 1080, 3, 64, 5.0))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/configs_r1.json", "w") as f:
    json.dump(out, f, indent=1)
