"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck):
device BVH build, device scene repack, generate -> bin -> extend -> accumulate (overlapped), shade, colour,
on the room with a reduced ray count, plus a tiny mesh with degenerate leaves."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
sim.set_device_bvh(True)
sim.load_mesh("testroomopt")
sim.init("route")
sim.set_params(maxIterations=2, photonCount=1 << 21)
dose = sim.run()
print("room dose mean", float(dose.mean()), "launches", sim.ctx.launch_count())
ctx = sim.ctx
rng = np.random.default_rng(1)
for n in (1, 5, 300, 70000):
    m = np.zeros((n, 16), dtype=np.float32)
    c = rng.uniform(-2, 2, (n, 3))
    for k in range(3):
        m[:, 4 * k:4 * k + 3] = (c + rng.uniform(-0.1, 0.1, (n, 3))).astype(np.float32)
    if n >= 5:
        m[1:4] = m[0]
    t, nodes, idx = ctx.build_bvh(m)
    ctx.upload_scene(t, nodes, idx)
    ctx.reset(True)
    ctx.trace_counts((0.0, 0.0, 0.0), 1.0, 0, 100000, 7)
    print(n, "tris: hits", int(ctx.read(uv.BUF.COUNTS).sum()))
sim.close()
print("SANITIZER_PROBE_DONE")
