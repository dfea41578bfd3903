import os, sys, time, json
sys.path.insert(0, os.getcwd())
import torch, bench
from torpedo_b200 import engine as E
g = bench.scene_cached(bench.N_GAUSSIANS)
scene = E.Scene(); scene.add_group(g)
w,h=bench.WIDTH,bench.HEIGHT
eng = E.GaussianEngine(w,h); eng.compile(scene, E.Settings(3))
cams=[]
for v in range(8):
    cam=E.PerspectiveCamera(w,h); cam.look_at(E.to_cartesian(*bench.ring_camera_params(v*8)),(0,0,0),(0,0,1)); cams.append(cam)
ubos=[c.pack() for c in cams]
for r in range(12): eng.raster_frame(cams[r%8])
eng.finish(); torch.cuda.synchronize()
K=256
t0=time.perf_counter()
for r in range(K): eng.raster_ubo(ubos[r%8], 3, torch.cuda.current_stream().cuda_stream)
t1=time.perf_counter()
eng.finish(); torch.cuda.synchronize()
t2=time.perf_counter()
print(json.dumps({"host_enqueue_ms_per_frame": (t1-t0)*1e3/K, "total_ms_per_frame": (t2-t0)*1e3/K}))
# from an idle GPU and an empty launch queue: what the host needs per frame when nothing pushes back
res=[]
for rep in range(5):
    torch.cuda.synchronize(); eng.finish()
    t0=time.perf_counter()
    for r in range(3): eng.raster_ubo(ubos[r%8], 3, torch.cuda.current_stream().cuda_stream)
    t1=time.perf_counter()
    eng.finish(); torch.cuda.synchronize()
    res.append((t1-t0)*1e3/3)
print(json.dumps({"host_enqueue_ms_per_frame_idle_queue": res}))
