"""Times the pieces of one end-to-end step (host wall clock, GPU box)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
sim.load_mesh("testroomopt"); sim.init("route")
tris, nodes, idx = sim.mesh_data()
ctx = sim.ctx
def T(name, fn, reps=5):
    ts = []
    for _ in range(reps):
        ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:24s} min {min(ts):9.3f} ms  med {sorted(ts)[len(ts)//2]:9.3f} ms", flush=True)
T("upload_scene", lambda: ctx.upload_scene(tris, nodes, idx))
T("set_params", lambda: sim.set_params(maxIterations=1))
T("reset_dosage_map", lambda: sim.reset_dosage_map())
def tick():
    sim.set_params(maxIterations=1); sim.reset_dosage_map(); sim.tick()
T("reset+tick", tick)
T("shade", lambda: sim.shade())
T("read_dose", lambda: sim.read_dose())
T("mesh_info", lambda: sim.mesh_info())
T("sim.ctx", lambda: sim.ctx)

import numpy as np
def e2e_step(log=False):
    ts = [time.perf_counter()]
    ctx.upload_scene(tris, nodes, idx); ts.append(time.perf_counter())
    sim.set_params(maxIterations=1); sim.reset_dosage_map(); ts.append(time.perf_counter())
    fin = False
    n = 0
    while not fin:
        fin = sim.tick(); n += 1
    ts.append(time.perf_counter())
    sim.reduce(); sim.shade(); ts.append(time.perf_counter())
    d = sim.read_dose(); ts.append(time.perf_counter())
    if log:
        print("e2e pieces ms:", [round((b - a) * 1e3, 3) for a, b in zip(ts, ts[1:])], "ticks", n, flush=True)
    return d
# first a long run like bench.py does
sim.set_params(maxIterations=20); sim.reset_dosage_map()
for _ in range(20): sim.tick()
for _ in range(4): e2e_step(True)
