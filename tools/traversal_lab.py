"""Driver of tools/traversal_lab.c: traversal statistics and the certified fast traversal on the host.

    python tools/traversal_lab.py [--scene room|soup1m|soup10m] [--rays 2000000] [--route route] [--positions 0,5,11]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import uvrt_testlib as T  # noqa: E402


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("rays", "innerVisits", "leafVisits", "triTests", "hits", "deadInner", "deadLeaf", "pops",
                                          "accepted", "inverted")] + [("maxInvAbs", C.c_double), ("maxInvRel", C.c_double)] + \
               [(k, C.c_uint64) for k in ("certFail", "mismatch", "rawMismatch", "nearTie", "boxReject", "tminFail", "warpIters", "warpLaneIters")]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        r = max(1, d["rays"])
        d["inner_per_ray"] = d["innerVisits"] / r
        d["tri_per_ray"] = d["triTests"] / r
        d["leaf_per_ray"] = d["leafVisits"] / r
        d["simt_eff"] = d["warpLaneIters"] / max(1, 32 * d["warpIters"])
        return d


def lab():
    src = os.path.join(ROOT, "tools", "traversal_lab.c")
    os.makedirs(os.path.join(ROOT, "tools", "_build"), exist_ok=True)
    so = os.path.join(ROOT, "tools", "_build", "liblab.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so, "-lm"], check=True)
    L = C.CDLL(so)
    vp = C.c_void_p
    L.lab_exact_stats.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int, C.POINTER(Stats)]
    L.lab_fast.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.POINTER(Stats)]
    return L


def bin_order(rays, y0, length, nY=16, nT=32, nP=128):
    d, o = rays["dir"], rays["orig"]
    t = np.clip(((d[:, 1] + 1) * 0.5 * nT).astype(np.int64), 0, nT - 1)
    ph = np.clip(((np.arctan2(d[:, 2], d[:, 0]) + np.pi) * 0.159154943 * nP).astype(np.int64), 0, nP - 1)
    y = np.clip(((o[:, 1] - y0) / length * nY).astype(np.int64), 0, nY - 1)
    return np.argsort((t * nY + y) * nP + ph, kind="stable")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="room")
    ap.add_argument("--rays", type=int, default=1_000_000)
    ap.add_argument("--route", default="route")
    ap.add_argument("--positions", default="0,5,11")
    ap.add_argument("--drel", type=float, default=2.0 ** -12)
    ap.add_argument("--dabs", type=float, default=2.0 ** -14)
    args = ap.parse_args()
    uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
    B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
    f32 = np.float32
    if args.scene == "room":
        sim = uv.Sim(asset_root=T.DATA)
        sim.load_mesh("testroomopt")
        sim.load_route(args.route)
        tris, nodes, tri_idx = sim.mesh_data()
        floor = sim.mesh_info()["floor"]
        p = sim.params
        lamps = [(f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y)) for x, y, _ in sim.positions]
        length = p.lightLength
        sim.close()
    else:
        from soup import make_soup, soup_route
        n = {"soup1m": 1_000_000, "soup10m": 10_000_000, "soup200k": 200_000}[args.scene]
        tris, nodes, tri_idx = B.build_bvh(make_soup(n))
        lamps = [(f32(x), f32(0.5), f32(z)) for x, z, _ in soup_route()]
        length = 1.0
    L, O = lab(), T.oracle()
    for k in [int(x) for x in args.positions.split(",")]:
        lp = lamps[k]
        rays = np.zeros(args.rays, dtype=T.RAY_DT)
        O.orc_generate(T.ptr(rays), 0, args.rays, lp[0], lp[1], lp[2], f32(length), 7 * k, None)
        rays = rays[bin_order(rays, lp[1], length)].copy()
        ref = rays.copy()
        s = Stats()
        L.lab_exact_stats(T.ptr(tris), T.ptr(ref), T.ptr(nodes), T.ptr(tri_idx), args.rays, 0, C.byref(s))
        e = s.as_dict()
        print(json.dumps({"pos": k, "mode": "exact", **{k2: e[k2] for k2 in ("inner_per_ray", "leaf_per_ray", "tri_per_ray", "pops", "deadInner", "deadLeaf",
                                                                               "accepted", "inverted", "maxInvAbs", "maxInvRel", "simt_eff", "warpIters")}}))
        ref2 = rays.copy()
        s2 = Stats()
        L.lab_exact_stats(T.ptr(tris), T.ptr(ref2), T.ptr(nodes), T.ptr(tri_idx), args.rays, 1, C.byref(s2))
        assert ref2.tobytes() == ref.tobytes(), "pop check on inner nodes changed a result"
        e2 = s2.as_dict()
        print(json.dumps({"pos": k, "mode": "exact+popcheck", "inner_per_ray": e2["inner_per_ray"], "simt_eff": e2["simt_eff"], "warpIters": e2["warpIters"]}))
        for quant, popc in ((0, 0), (0, 1), (1, 1)):
            out = rays.copy()
            s3 = Stats()
            L.lab_fast(T.ptr(tris), T.ptr(out), T.ptr(ref), T.ptr(nodes), len(nodes), T.ptr(tri_idx), args.rays, quant,
                       args.drel, args.dabs, popc, C.byref(s3))
            e3 = s3.as_dict()
            print(json.dumps({"pos": k, "mode": f"fast quant={quant} popcheck={popc}", **{k2: e3[k2] for k2 in (
                "inner_per_ray", "leaf_per_ray", "tri_per_ray", "certFail", "nearTie", "tminFail", "boxReject", "mismatch", "rawMismatch",
                "simt_eff", "warpIters")}}))


if __name__ == "__main__":
    main()
