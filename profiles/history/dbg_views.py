import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from torpedo_b200 import engine as E, scenes
w, h = 256, 144
g = scenes.garden(20000, seed=61, log_scale_mean=-3.6)
ubos = []
for k in range(6):
    cam = E.PerspectiveCamera(w, h); cam.look_at(E.to_cartesian(2*np.pi*k/6, 0.9, 5.0), (0,0,0), (0,0,1)); ubos.append(cam.pack())
scene = E.Scene(); scene.add_group(g)
eng = E.GaussianEngine(w, h); eng.compile(scene)
single = []
for k in range(6):
    eng.raster_ubo(ubos[k], 3); single.append(eng.draw().copy())
eng.close()
for graph in (1, 0):
    scene = E.Scene(); scene.add_group(g)
    eng = E.GaussianEngine(w, h); eng.compile(scene); eng.graph_replay(graph)
    for trial in range(2):
        frames = torch.zeros((6, h, w, 4), dtype=torch.uint8, device='cuda')
        eng.raster_views(np.stack(ubos), frames.data_ptr(), h*w*4, 3, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out = frames.cpu().numpy()
        print('graph', graph, 'trial', trial, [int(np.abs(out[k].astype(int)-single[k].astype(int)).max()) for k in range(6)], eng.capacity(), eng.graph_replay())
    eng.close()
