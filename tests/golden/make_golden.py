"""Generates tests/golden/appendix_c.json from the reference's OWN sources compiled by
oracle/build_ref.sh (oracle/_ref/libuvrt_ref.so): bvh.cpp unmodified and cl/*.cl after the
vector-literal rewrite.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The values reproduce SURVEY.md App. C (launch semantics of App. B) and add a few raw samples
so the GPU box -- which has no /root/reference -- can check against the reference's behaviour.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import uvrt_testlib as T  # noqa: E402


def main():
    T.build_checkers()
    R = T.ref()
    tris0 = T.load_glb_tris(T.ROOM)
    floor = T.floor_height(tris0)
    tris, nodes, tri_idx = T.ref_build_bvh(tris0)
    pre = T.reachable_preorder(nodes)
    buf = bytearray()
    for i in pre:
        buf += np.uint32(i).tobytes() + nodes[i].tobytes()
    out = {
        "scene": {"triangles": int(tris.shape[0]), "floor_bits": int(np.float32(floor).view(np.uint32)),
                  "reachable_nodes": int(len(pre)), "max_node_index": int(pre.max()),
                  "fnv_triIdx": f"{T.fnv(tri_idx):016x}",
                  "fnv_nodes_preorder": f"{T.fnv(np.frombuffer(bytes(buf), dtype=np.uint8)):016x}",
                  "triIdx_head": [int(x) for x in tri_idx[:8]]},
        "launch": {},
    }
    f32 = np.float32
    lp = (f32(-0.25500134), f32(f32(floor) + f32(0.40000001)), f32(-3.3149862))
    out["lightPos_bits"] = [int(np.float32(v).view(np.uint32)) for v in lp]
    for P in (1000000, 2796202):
        rays = np.zeros(P, dtype=T.RAY_DT)
        so = C.c_uint(0)
        R.ref_generate(T.ptr(rays), 0, P, lp[0], lp[1], lp[2], 1.0, 0, C.byref(so))
        fnv_gen = T.fnv(rays)
        temp = np.zeros(tris.shape[0], dtype=np.int32)
        R.ref_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, tris.shape[0], 0)
        sample_idx = [0, 1, 2, 3, 1000, P - 1]
        out["launch"][str(P)] = {
            "seed_out": int(so.value), "fnv_rays_after_generate": f"{fnv_gen:016x}",
            "fnv_hits": f"{int(T.oracle().orc_fnv_hits(T.ptr(rays), P)):016x}",
            "fnv_counts": f"{T.fnv(temp):016x}", "hits": int(temp.sum()),
            "hottest": [int(temp.argmax()), int(temp.max())],
            "samples": {str(i): rays[i].tobytes().hex() for i in sample_idx},
        }
    # one full pass over lange_route (App. C.2): SEED chain, hits per position, dose
    import importlib
    uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
    sim = uv.Sim(asset_root="/root/reference")
    sim.load_route("lange_route")
    pos = sim.positions
    p = sim.params
    sim.close()
    P = int(p.photonsPerLight)
    n = tris.shape[0]
    photon = np.zeros(n); mx = np.zeros(n); temp = np.zeros(n, dtype=np.int32)
    seed = 0
    chain, hits = [0], []
    rays = np.zeros(P, dtype=T.RAY_DT)
    for (x, y, dur) in pos:
        so = C.c_uint(0)
        R.ref_generate(T.ptr(rays), 0, P, f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y), f32(p.lightLength), seed, C.byref(so))
        R.ref_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, n, 0)
        hits.append(int(temp.sum()))
        R.ref_accumulate(T.ptr(photon), T.ptr(mx), T.ptr(temp), f32(dur), n)
        seed = int(so.value)
        chain.append(seed)
    dose = np.zeros(n, dtype=np.float32)
    R.ref_compute_dosage(T.ptr(photon), T.ptr(dose), T.ptr(tris), P, f32(f32(p.lightIntensity) * f32(0.1)), n)
    color = np.zeros((n, 9), dtype=np.float32)
    R.ref_dosage_to_color(T.ptr(dose), T.ptr(color), f32(p.minDosage), 0, n)
    out["pass_lange_route"] = {
        "photonsPerLight": P, "seed_chain": chain, "hits_per_position": hits,
        "fnv_photonMap": f"{T.fnv(photon):016x}", "fnv_maxPhotonMap": f"{T.fnv(mx):016x}",
        "fnv_dose": f"{T.fnv(dose):016x}", "fnv_color": f"{T.fnv(color):016x}",
        "dose_head_bits": [int(v) for v in dose[:8].view(np.uint32)],
        "dose_max": float(dose.max()), "dose_mean": float(dose.astype(np.float64).mean()),
        "unlit": int((dose == 0).sum()),
    }
    json.dump(out, open(os.path.join(HERE, "appendix_c.json"), "w"), indent=1)
    print(json.dumps(out["pass_lange_route"], indent=1)[:1200])


if __name__ == "__main__":
    main()
