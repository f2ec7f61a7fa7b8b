"""Incoherent-traversal stress (BASELINE.json config 5): synthetic triangle soup, ray sweep.

    python tools/soup_bench.py --tris 1000000 --rays 1000000,10000000

(The comparison with the oracle lives in tests/test_gpu_parity.py::test_soup_scene_config5; set
UVRT_SOUP_TRIS=10000000 to run it at full size.)
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
from soup import make_soup, soup_route  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--size", type=float, default=0.01)
    ap.add_argument("--rays", default="1000000")
    args = ap.parse_args()
    t0 = time.perf_counter()
    tris = make_soup(args.tris, args.size)
    t1 = time.perf_counter()
    sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
    sim.set_triangles(tris)
    t2 = time.perf_counter()
    sim.set_positions(soup_route())
    sim.set_params(lightHeight=0.5, lightLength=1.0, lightIntensity=100.0, maxIterations=1, photonCount=1 << 20)
    sim.init(None)
    ctx = sim.ctx
    info = ctx.scene_info()
    print(json.dumps({"tris": args.tris, "make_s": round(t1 - t0, 2), "bvh_build_s": round(t2 - t1, 2), "scene": info,
                      "floor": float(sim.mesh_info()["floor"]), "tame": ctx.get_option("scene_tame")}), flush=True)
    tb = time.perf_counter()
    dt, dn, di = ctx.build_bvh(tris)
    tb = time.perf_counter() - tb
    ht, hn, hi = sim.mesh_data()
    print(json.dumps({"device_bvh_build_s": round(tb, 3), "same_triIdx": bool(np.array_equal(di, hi)),
                      "same_nodes": bool(dn.tobytes() == hn.tobytes())}), flush=True)
    floor = sim.mesh_info()["floor"]
    pos = sim.positions
    for P in [int(x) for x in args.rays.split(",")]:
        lp = (np.float32(pos[5, 0]), np.float32(np.float32(floor) + np.float32(0.5)), np.float32(pos[5, 1]))
        for binned in (0, 1):
            ctx.set_option("bin_rays", binned)
            times = []
            for r in range(4):
                ctx.reset(False)
                ctx.mark(0)
                ctx.trace_counts(lp, 1.0, 0, P, 0)
                ctx.mark(1)
                times.append(ctx.elapsed_ms(0, 1))
            counts = ctx.read(uv.BUF.COUNTS)
            best = min(times[1:])
            print(json.dumps({"rays": P, "bin_rays": binned, "ms": round(best, 3), "mrays_s": round(P / best / 1e3, 1),
                              "hits": int(counts.sum())}), flush=True)


if __name__ == "__main__":
    main()
