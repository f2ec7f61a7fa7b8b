"""Generates tests/golden/route_runs.json and tests/golden/soup.json from the reference's OWN sources
compiled by oracle/build_ref.sh (oracle/_ref/libuvrt_ref.so: bvh.cpp unmodified, cl/*.cl after the
vector-literal rewrite).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_runs.py [--skip-soup10m]

route_runs.json -- BASELINE configs[1] exactly as RayTracer drives it (raytracer.cpp:66-120): the shipped
    positions/route.xml (and lange_route.xml as the config-2 substitute), 2^25 photons, 10 iterations =
    120 launches of 2,796,202 rays: SEED chain, hits per launch, FNV of every launch's count vector, and after
    every iteration the FNVs of the f64 photon map / max map and of the f32 dose + colours RayTracer::Shade
    would produce at that point (dose divisor photonMapSize / L, raytracer.cpp:111).
soup.json -- BASELINE config 5: the synthetic triangle soup of tools/soup.py (1 M and 10 M triangles) through the
    reference's builder and extend kernel: FNVs of triIdx, of the (dist, triID) pairs and of the count vector of
    1e6-ray launches.

About 6 minutes on 8 cores (335 M rays per route at ~2.3 Mrays/s, 10 M-triangle build + 2 x 1e6 incoherent rays).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "tools"))
import uvrt_testlib as T  # noqa: E402

f32 = np.float32


def route_run(R, tris, nodes, tri_idx, floor, route, iterations=10):
    uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
    sim = uv.Sim(asset_root="/root/reference")
    sim.load_route(route)
    pos, p = sim.positions, sim.params
    sim.close()
    P, L, n = int(p.photonsPerLight), len(pos), tris.shape[0]
    photon, mx, temp = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    rays = np.zeros(P, dtype=T.RAY_DT)
    seed, chain, hits, fnv_counts, per_iter = 0, [0], [], [], []
    for it in range(iterations):
        for (x, y, dur) in pos:
            so = C.c_uint(0)
            R.ref_generate(T.ptr(rays), 0, P, f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y), f32(p.lightLength), seed, C.byref(so))
            R.ref_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, n, 0)
            hits.append(int(temp.sum()))
            fnv_counts.append(f"{T.fnv(temp):016x}")
            R.ref_accumulate(T.ptr(photon), T.ptr(mx), T.ptr(temp), f32(dur), n)
            seed = int(so.value)
            chain.append(seed)
        # RayTracer::Shade at this point of the run (raytracer.cpp:104-119): divisor photonMapSize / L
        per_light = ((it + 1) * L * P) // L
        dose = np.zeros(n, dtype=np.float32)
        R.ref_compute_dosage(T.ptr(photon), T.ptr(dose), T.ptr(tris), per_light, f32(f32(p.lightIntensity) * f32(0.1)), n)
        color = np.zeros((n, 9), dtype=np.float32)
        R.ref_dosage_to_color(T.ptr(dose), T.ptr(color), f32(p.minDosage), 0, n)
        per_iter.append({"fnv_photonMap": f"{T.fnv(photon):016x}", "fnv_maxPhotonMap": f"{T.fnv(mx):016x}",
                         "fnv_dose": f"{T.fnv(dose):016x}", "fnv_color": f"{T.fnv(color):016x}", "seed": seed,
                         "dose_head_bits": [int(v) for v in dose[:8].view(np.uint32)],
                         "dose_mean": float(dose.astype(np.float64).mean()), "dose_max": float(dose.max())})
        print(f"  {route}: iteration {it + 1}/{iterations} dose fnv {per_iter[-1]['fnv_dose']}", flush=True)
    return {"route": route + ".xml", "photonsPerLight": P, "positions": L, "iterations": iterations,
            "rays": iterations * L * P, "lightHeight_bits": int(f32(p.lightHeight).view(np.uint32)),
            "lightIntensity_bits": int(f32(p.lightIntensity).view(np.uint32)),
            "seed_chain": chain, "hits_per_launch": hits, "fnv_counts_per_launch": fnv_counts, "after_iteration": per_iter}


def soup_case(R, O, n_tris, launches, P=1_000_000):
    from soup import make_soup, soup_route
    t0 = time.time()
    tris, nodes, tri_idx = T.ref_build_bvh(make_soup(n_tris))
    print(f"  soup {n_tris}: reference builder {time.time() - t0:.1f} s, {len(nodes)} node slots", flush=True)
    pre = T.reachable_preorder(nodes) if n_tris <= 1_000_000 else None
    out = {"triangles": n_tris, "nodesUsed": int(len(nodes)), "fnv_triIdx": f"{T.fnv(tri_idx):016x}", "launch": {}}
    if pre is not None:
        out["reachable_nodes"] = int(len(pre))
    route = soup_route()
    for (j, seed) in launches:
        x, z, _ = route[j]
        lp = (f32(x), f32(0.5), f32(z))
        # the port first: the compiled reference never returns from a work-item whose RNG state is 0
        rays = np.zeros(P, dtype=T.RAY_DT)
        so = C.c_uint32(0)
        O.orc_generate(T.ptr(rays), 0, P, lp[0], lp[1], lp[2], f32(1.0), seed, C.byref(so))
        stuck = np.flatnonzero((rays["dir"][:, 1] == f32(-1.0)) & (rays["dir"][:, 0] == 0) & (rays["dir"][:, 2] == 0))
        use_ref = stuck.size == 0
        if use_ref:
            rr = np.zeros(P, dtype=T.RAY_DT)
            so2 = C.c_uint(0)
            R.ref_generate(T.ptr(rr), 0, P, lp[0], lp[1], lp[2], f32(1.0), seed, C.byref(so2))
            assert rr.tobytes() == rays.tobytes() and so2.value == so.value
        temp = np.zeros(n_tris, dtype=np.int32)
        t0 = time.time()
        if use_ref:
            R.ref_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, n_tris, 0)
        else:
            O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, 0, None)
        print(f"  soup {n_tris}: launch {j} seed {seed}: {P / (time.time() - t0) / 1e6:.2f} Mrays/s "
              f"({'reference' if use_ref else 'port (zero RNG state in this launch)'})", flush=True)
        out["launch"][f"{j}:{seed}"] = {
            "position": j, "seed_in": seed, "rays": P, "seed_out": int(so.value), "source": "reference" if use_ref else "port",
            "fnv_hits": f"{int(O.orc_fnv_hits(T.ptr(rays), P)):016x}", "fnv_counts": f"{T.fnv(temp):016x}", "hits": int(temp.sum()),
            "mean_hit_dist": float(rays["dist"][rays["dist"] != f32(1e30)].astype(np.float64).mean()),
        }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-soup10m", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    T.build_checkers()
    R, O = T.ref(), T.oracle()
    if args.only in ("", "routes"):
        tris0 = T.load_glb_tris(T.ROOM)
        floor = T.floor_height(tris0)
        tris, nodes, tri_idx = T.ref_build_bvh(tris0)
        out = {"generator": "tests/golden/make_golden_runs.py (oracle/_ref = the reference's own bvh.cpp + cl/*.cl)",
               "room": "testroomopt.glb", "triangles": int(tris.shape[0]), "runs": {}}
        for route in ("route", "lange_route"):
            t0 = time.time()
            out["runs"][route] = route_run(R, tris, nodes, tri_idx, floor, route)
            print(f"{route}: {time.time() - t0:.0f} s", flush=True)
        json.dump(out, open(os.path.join(HERE, "route_runs.json"), "w"), indent=1)
    if args.only in ("", "soup"):
        soup = {"generator": "tests/golden/make_golden_runs.py; scene = tools/soup.py (SURVEY 8d config 5)", "scenes": {}}
        soup["scenes"]["1000000"] = soup_case(R, O, 1_000_000, [(5, 3), (0, 0)])
        if not args.skip_soup10m:
            soup["scenes"]["10000000"] = soup_case(R, O, 10_000_000, [(0, 0), (5, 3)])
        json.dump(soup, open(os.path.join(HERE, "soup.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
