"""BVH build: device (uvrt_build_bvh) against the host builder (host/bvh.cpp = the reference's bvh.cpp),
on the room and on triangle soups.  Prints one JSON line per mesh; the trees are compared bit for bit.

    python tools/bvh_build_bench.py [--soup 1000000,10000000] [--reps 3]
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
from soup import make_soup  # noqa: E402


def bench(ctx, name, tris, reps):
    t0 = time.perf_counter()
    ht, hn, hi = B.build_bvh(tris)
    host_s = time.perf_counter() - t0
    times = []
    n = tris.shape[0]
    out = (np.empty_like(tris), np.empty(2 * n + 64, dtype=B.NODE_DTYPE), np.empty(n, dtype=np.uint32))
    for a in out:
        a.view(np.uint8).fill(0)                  # touch the pages: no page faults inside the timed copies
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        dt, dn, di = ctx.build_bvh(tris, out)
        times.append(time.perf_counter() - t0)
    same = bool(np.array_equal(di, hi) and dn.tobytes() == hn.tobytes() and dt.tobytes() == ht.tobytes())
    print(json.dumps({"mesh": name, "triangles": int(tris.shape[0]), "nodes_used": int(len(dn)),
                      "host_build_ms": round(host_s * 1e3, 2), "host_threads": os.cpu_count(),
                      "device_build_first_ms": round(times[0] * 1e3, 2),
                      "device_build_ms": round(min(times[1:] or times) * 1e3, 2),
                      "what": "uvrt_build_bvh wall clock: H2D of the triangles, build, D2H of nodes + triIdx + centroids",
                      "identical_to_host_tree": same}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--soup", default="1000000,10000000")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
    sim.load_mesh("testroomopt")
    tris = sim.mesh_data()[0].copy()
    ctx = uv.Context(0)
    bench(ctx, "testroomopt.glb", tris, args.reps)
    for n in [int(x) for x in args.soup.split(",") if x]:
        bench(ctx, f"soup {n}", make_soup(n), args.reps)
    ctx.close()


if __name__ == "__main__":
    main()
