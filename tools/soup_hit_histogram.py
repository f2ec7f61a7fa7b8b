#!/usr/bin/env python3
"""Hit-distance histogram of BASELINE config 5 ("the builder reports the measured
hit-distance histogram"): the 10 M-triangle soup, 12 lamp positions, traced on the host by the checker
(oracle/uvrt_oracle.c), with the traversal counters of the reference's tree.  CPU only.

    python tools/soup_hit_histogram.py [--triangles 10000000] [--rays 200000] > profiles/r2_soup_hit_histogram.json
"""
import argparse, ctypes as C, importlib, json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import uvrt_testlib as T
from soup import make_soup, soup_route


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--triangles", type=int, default=10_000_000)
    ap.add_argument("--rays", type=int, default=200_000, help="rays per lamp position")
    a = ap.parse_args()
    B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
    O = T.oracle()
    f32 = np.float32
    tris, nodes, ti = B.build_bvh(make_soup(a.triangles))
    edges = np.array([0, 0.05, 0.1, 0.25, 0.5, 1, 1.5, 2, 3, 4, 6, 8, 12], dtype=np.float64)
    hist = np.zeros(len(edges) - 1, dtype=np.int64)
    tot = T.Counters()
    dists = []
    for k, (x, z, _) in enumerate(soup_route()):
        rays = np.zeros(a.rays, dtype=T.RAY_DT)
        O.orc_generate(T.ptr(rays), 0, a.rays, f32(x), f32(0.5), f32(z), f32(1.0), 3 + 977 * k, None)
        temp = np.zeros(a.triangles, dtype=np.int32)
        c = T.Counters()
        O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(ti), a.rays, 1, C.byref(c))
        for f, _t in T.Counters._fields_:
            setattr(tot, f, max(getattr(tot, f), getattr(c, f)) if f == "maxStack" else getattr(tot, f) + getattr(c, f))
        d = rays["dist"][rays["dist"] != f32(1e30)].astype(np.float64)
        dists.append(d)
        hist += np.histogram(d, bins=edges)[0]
    d = np.concatenate(dists)
    n = int(tot.rays)
    print(json.dumps({
        "scene": f"soup of {a.triangles} triangles (tools/soup.py), 12 lamp positions x {a.rays} rays, checker on the host",
        "rays": n, "hits": int(tot.hits), "hit_fraction": round(tot.hits / n, 4),
        "inner_visits_per_ray": round(tot.innerVisits / n, 2), "leaf_visits_per_ray": round(tot.leafVisits / n, 2),
        "triangle_tests_per_ray": round(tot.triTests / n, 2), "max_stack": int(tot.maxStack),
        "hit_distance_m": {"mean": round(float(d.mean()), 3), "median": round(float(np.median(d)), 3),
                           "p90": round(float(np.quantile(d, 0.9)), 3), "max": round(float(d.max()), 3)},
        "histogram_edges_m": edges.tolist(), "histogram_hits": hist.tolist(),
    }, indent=1))


if __name__ == "__main__":
    main()
