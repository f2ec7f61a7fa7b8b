"""The certified fast extend (variant 50) against the exact kernel (variant 2): time, results, fallbacks.

    python tools/fast_sweep.py [--scene room|soup1m|soup10m] [--rays N] [--positions 0,5,11] [--cfgs 0,1] [--chunks 64] [--check]

Per lamp position: one launch generated on the device, then bin + extend timed (CUDA events, best of 5) for every
variant / register budget; the ray records (dist bits, triID) and the count vector must be byte-identical to variant 2's.
--check repeats every fast launch with "fast_check" = 1 (every ray traced both ways) and prints the number of
CERTIFIED rays whose answer differed from the reference-order traversal (the scheme's claim is that this is 0).
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="room")
    ap.add_argument("--rays", type=int, default=2796202)
    ap.add_argument("--positions", default="0,5,11")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--variants", default="2,50")
    ap.add_argument("--cfgs", default="0,1", help='"fast_cfg" values of variant 50: 0 / 1 register budgets, 2 = refill kernel (make EXPERIMENTS=1 builds only)')
    ap.add_argument("--chunks", default="64", help='"refill_chunk" values (fast_cfg 2 only)')
    args = ap.parse_args()
    f32 = np.float32
    if args.scene == "room":
        sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
        sim.load_mesh("testroomopt")
        sim.init("route")
        c = sim.ctx
        floor = sim.mesh_info()["floor"]
        p = sim.params
        lamps = [(f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y)) for x, y, _ in sim.positions]
        length = p.lightLength
    else:
        from soup import make_soup, soup_route
        n = {"soup1m": 1_000_000, "soup10m": 10_000_000, "soup200k": 200_000}[args.scene]
        c = uv.Context(0)
        tris, nodes, tri_idx = c.build_bvh(make_soup(n))
        c.upload_scene(tris, nodes, tri_idx)
        del tris, nodes, tri_idx
        lamps = [(f32(x), f32(0.5), f32(z)) for x, z, _ in soup_route()]
        length = 1.0
    print(json.dumps({"scene": args.scene, "info": c.scene_info(), "tame": c.get_option("scene_tame"), "nested": c.get_option("scene_nested"),
                      "fast_ready": c.get_option("fast_ready")}), flush=True)
    P = args.rays
    for pi in [int(x) for x in args.positions.split(",")]:
        lp = lamps[pi]
        ref = None
        for v in [int(x) for x in args.variants.split(",")]:
            cfgs = [(int(x), 0) for x in args.cfgs.split(",") if int(x) != 2] if v >= 50 else [(0, 0)]
            if v >= 50 and "2" in args.cfgs.split(","):
                cfgs += [(2, int(x)) for x in args.chunks.split(",")]
            for cfg, chunk in cfgs:
                c.set_option("extend_variant", v)
                c.set_option("fast_cfg", cfg)
                if chunk:
                    c.set_option("refill_chunk", chunk)
                c.set_option("fast_check", 0)
                c.fast_stats(reset=True)
                times = []
                for r in range(args.reps + 2):
                    c.reset(False)
                    c.generate(lp, length, 0, P, 7 * pi)
                    c.mark(0)
                    c.extend(P)
                    c.mark(1)
                    t = c.elapsed_ms(0, 1)
                    if r >= 2:
                        times.append(t)
                rays = c.read(uv.BUF.RAYS, P)
                counts = c.read(uv.BUF.COUNTS)
                st = c.fast_stats(reset=True)
                if ref is None:
                    ref = (rays, counts)
                bad = int(np.count_nonzero((rays["dist"].view(np.uint32) != ref[0]["dist"].view(np.uint32)) | (rays["triID"] != ref[0]["triID"])))
                out = {"scene": args.scene, "pos": pi, "variant": v, "fast_cfg": cfg, "refill_chunk": chunk, "ms_best": round(min(times), 4), "ms_med": round(float(np.median(times)), 4),
                       "mrays_s": round(P / min(times) / 1e3, 1), "rays_differ": bad, "counts_equal": bool(np.array_equal(counts, ref[1])),
                       "cert_fallbacks_per_launch": st["cert_fallbacks"] // (args.reps + 2), "ineligible_per_launch": st["ineligible"] // (args.reps + 2)}
                if args.check and v >= 50 and cfg in (0, 2):
                    c.set_option("fast_check", 1)
                    c.reset(False)
                    c.generate(lp, length, 0, P, 7 * pi)
                    c.extend(P)
                    out["check_mismatches"] = c.fast_stats(reset=True)["check_mismatches"]
                    c.set_option("fast_check", 0)
                print(json.dumps(out), flush=True)
    c.set_option("extend_variant", -1)


if __name__ == "__main__":
    main()
