import os, sys, json
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
for fif in (1,2,3,4):
    eng.set_frames_in_flight(fif)
    for r in range(12): eng.raster_frame(cams[r%8])
    eng.finish()
    res=[]
    for rep in range(3):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        K=64
        for r in range(K): eng.raster_frame(cams[r%8])
        e1.record(); eng.finish(); torch.cuda.synchronize()
        res.append(round(e0.elapsed_time(e1)/K,4))
    print(json.dumps({"fif":fif,"ms_per_frame":res}))
