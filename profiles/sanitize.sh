# compute-sanitizer over the frame kernels (run under gpurun): memcheck, racecheck, synccheck on smoke(); memcheck over the edge cases
# usage: bash profiles/sanitize.sh > gpurun_out/sanitizer.txt
S=/usr/local/cuda/bin/compute-sanitizer
echo "== __graft_entry__.smoke() (20 k Gaussians, 256x144, all frame kernels) =="
for tool in memcheck racecheck synccheck; do
  echo -n "$tool: "; timeout 600 $S --tool $tool python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1
done
echo "== the words sort alone on 6 M tile-sort-like words (732 tiles: every persistent CTA loops over several, every look-back chain in use): racecheck, memcheck =="
for tool in racecheck memcheck; do
  echo -n "$tool: "; timeout 1500 $S --tool $tool profiles/micro/bin/os_trace 6000000 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sorted" | cut -c1-90 | tr '\n' ' '; echo
done
echo "== a 300 k-Gaussian frame at 640x360 (all frame kernels, 78 sort tiles): memcheck, racecheck =="
cat > /tmp/san_frame.py <<'P'
import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from oracle import oracle as O
from torpedo_b200 import engine as E, scenes
w, h = 640, 360
g = scenes.garden(300000, seed=7, log_scale_mean=-4.4)
sc = E.Scene(); sc.add_group(g)
eng = E.GaussianEngine(w, h); eng.compile(sc, E.Settings(3))
cam = E.PerspectiveCamera(w, h); cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
for _ in range(2):
    eng.raster_frame(cam)
img = eng.draw(); k, v = eng.read_sorted(); r = eng.read_ranges()
ref = O.render(g, cam.pack(), w, h, 3)
assert (k == ref.keys).all() and (v == ref.vals).all() and (r == ref.ranges).all()
assert np.abs(img.astype(int) - ref.rgba.astype(int)).max() <= 1
print("frame ok", len(k))
P
for tool in memcheck racecheck; do
  echo -n "$tool: "; timeout 1200 $S --tool $tool python /tmp/san_frame.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|frame ok|Error|error" | tail -2 | tr '\n' ' '; echo
done
echo "== memcheck over pytest -m gpu edge cases =="
timeout 1500 $S --tool memcheck python -m pytest tests/test_parity_gpu.py -q -k "edge or huge or capacity or entity or depth or in_flight or graph or degree or emitted or ply" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -2
