"""Times uvrt_upload_scene (host wall clock, synchronised) with the device repack and with the host repack."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
from soup import make_soup
sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
sim.load_mesh("testroomopt")
scenes = [("testroomopt.glb", sim.mesh_data())]
ctx = uv.Context(0)
for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1000000").split(",") if x]:
    scenes.append((f"soup {n}", ctx.build_bvh(make_soup(n))))
for name, (tris, nodes, idx) in scenes:
    row = {"scene": name, "triangles": int(tris.shape[0]), "nodes": int(len(nodes))}
    for mode, key in ((0, "device_repack_ms"), (1, "host_repack_ms")):
        ctx.set_option("host_repack", mode)
        ts = []
        for _ in range(6):
            ctx.sync(); t0 = time.perf_counter(); ctx.upload_scene(tris, nodes, idx); ts.append((time.perf_counter() - t0) * 1e3)
        row[key] = round(min(ts[1:]), 3)
        row[key.replace("_ms", "_bytes")] = ctx.scene_upload_bytes()
    print(json.dumps(row), flush=True)
