"""Times extend variants on one launch of the default route (2,796,202 rays, lange_route position 0
on testroomopt) with CUDA events on the context's stream.  Run on the GPU box:

    python tools/variant_sweep.py [--rays N] [--reps R] [--variants 0,1,2,...] [--bin "0;16,32,128"]

The timed region is uvrt_extend: (bin scan + scatter when binning is on) + the extend kernel.
Every configuration's per-triangle counts are compared with the first one's.
"""
import argparse
import importlib
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")


def part1by1(x):
    x = x.astype(np.uint32) & 0xffff
    x = (x | (x << 8)) & 0x00FF00FF
    x = (x | (x << 4)) & 0x0F0F0F0F
    x = (x | (x << 2)) & 0x33333333
    x = (x | (x << 1)) & 0x55555555
    return x


def sorted_rays(gen, spec, lp, length):
    """Host-side ray orderings (to explore what re-ordering can buy): equal-probability cells in
    (dir.y, azimuth) x origin slices.  spec = "none" or "nT,nP,nY[,mode]"."""
    if spec == "none":
        return gen
    parts = spec.split(",")
    nT, nP, nY = int(parts[0]), int(parts[1]), int(parts[2])
    mode = parts[3] if len(parts) > 3 else "tpy"
    d, o = gen["dir"], gen["orig"]
    t = np.clip(((d[:, 1] + 1.0) * 0.5 * nT).astype(np.int64), 0, nT - 1)
    ph = np.clip(((np.arctan2(d[:, 2], d[:, 0]) + np.pi) / (2 * np.pi) * nP).astype(np.int64), 0, nP - 1)
    y = np.clip(((o[:, 1] - lp[1]) / length * nY).astype(np.int64), 0, nY - 1)
    if mode == "morton":
        key = (part1by1(t) | (part1by1(ph) << 1)).astype(np.int64) * nY + y
    elif mode == "ytp":
        key = (y * nT + t) * nP + ph
    elif mode == "ytps":      # serpentine azimuth
        key = (y * nT + t) * nP + np.where(t % 2 == 0, ph, nP - 1 - ph)
    elif mode == "morton3":
        def p3(x):
            x = x.astype(np.uint64) & 0x3ff
            x = (x | (x << 16)) & 0x30000ff
            x = (x | (x << 8)) & 0x300f00f
            x = (x | (x << 4)) & 0x30c30c3
            x = (x | (x << 2)) & 0x9249249
            return x
        key = (p3(t) | (p3(ph) << 1) | (p3(y) << 2)).astype(np.int64)
    elif mode == "tyq":       # dir.y cell, then origin slice, then azimuth (what the device binning uses)
        key = (t * nY + y) * nP + ph
    else:
        key = (t * nP + ph) * nY + y
    return gen[np.argsort(key, kind="stable")]


def ints(s):
    return [int(x) for x in s.split(",")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=2796202)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variants", default="0,1,2")
    ap.add_argument("--hist", default="0", help="hist_mode values (persistent variants only)")
    ap.add_argument("--refill", default="24", help="refill thresholds (persistent variants only)")
    ap.add_argument("--chunk", default="128", help="rays per warp (chunk-persistent variants 40-43 only)")
    ap.add_argument("--generic", default="0", help="generic_octant values (one-thread-per-ray variants only)")
    ap.add_argument("--cfg", default="1", help="simple_cfg values (one-thread-per-ray variants only)")
    ap.add_argument("--fetch", default="0", help="fetch_mode values: 0 LSU, 1 texture, 2 mixed (variant 2 only)")
    ap.add_argument("--bin", default="16,32,128", help="device binning: 0 (off) or nY,nT,nP; ';'-separated")
    ap.add_argument("--sort", default="none", help="host-side orderings, ';'-separated (see sorted_rays)")
    ap.add_argument("--positions", default="0")
    ap.add_argument("--opt", action="append", default=[], help="extra backend option key=value, repeatable")
    args = ap.parse_args()

    sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
    sim.load_mesh("testroomopt")
    sim.init("lange_route")
    c = sim.ctx
    floor = sim.mesh_info()["floor"]
    p, pos, P = sim.params, sim.positions, args.rays
    for kv in args.opt:
        k, v = kv.split("=")
        c.set_option(k, int(v))
    print(json.dumps({"device": c.device_info(), "scene": c.scene_info()}))
    for pi in ints(args.positions):
        lp = (np.float32(pos[pi, 0]), np.float32(np.float32(floor) + np.float32(p.lightHeight)), np.float32(pos[pi, 1]))
        c.generate(lp, p.lightLength, 0, P, 0)
        gen0 = c.read(uv.BUF.RAYS, P)
        ref_counts = None
        for sort, binspec, v in itertools.product(args.sort.split(";"), args.bin.split(";"), ints(args.variants)):
            gen = sorted_rays(gen0, sort, lp, p.lightLength)
            if binspec == "0":
                c.set_option("bin_rays", 0)
            else:
                by, bt, bp = ints(binspec)
                c.set_option("bin_rays", 1)
                c.set_option("bin_y", by)
                c.set_option("bin_t", bt)
                c.set_option("bin_p", bp)
            persistent = v >= 10
            knobs = itertools.product(ints(args.hist) if persistent and v < 40 else [0],
                                      ints(args.refill) if persistent else [24],
                                      [1] if persistent else ints(args.cfg),
                                      ints(args.fetch) if v == 2 else [0],
                                      ints(args.chunk) if v >= 40 else [128],
                                      [0] if persistent else ints(args.generic))
            for hist, refill, cfg, fetch, chunk, generic in knobs:
                c.set_option("chunk", chunk)
                c.set_option("generic_octant", generic)
                c.set_option("extend_variant", v)
                c.set_option("hist_mode", hist)
                c.set_option("refill", refill)
                c.set_option("simple_cfg", cfg)
                c.set_option("fetch_mode", fetch)
                times = []
                for r in range(args.reps + 2):
                    c.reset(False)
                    if sort == "none":
                        c.generate(lp, p.lightLength, 0, P, 0)
                    else:
                        c.write(uv.BUF.RAYS, gen)
                    c.mark(0)
                    c.extend(P)
                    c.mark(1)
                    t = c.elapsed_ms(0, 1)
                    if r >= 2:
                        times.append(t)
                counts = c.read(uv.BUF.COUNTS)
                if ref_counts is None:
                    ref_counts = counts
                best = min(times)
                print(json.dumps({"pos": pi, "sort": sort, "bin": binspec, "variant": v, "hist": hist, "refill": refill,
                                  "cfg": cfg, "fetch": fetch, "chunk": chunk, "generic": generic, "ms_best": round(best, 4),
                                  "ms_med": round(float(np.median(times)), 4), "mrays_s": round(P / best / 1e3, 1),
                                  "counts_equal": bool(np.array_equal(counts, ref_counts))}), flush=True)
    for k, v in (("extend_variant", -1), ("bin_rays", 1), ("hist_mode", 0), ("fetch_mode", 3), ("simple_cfg", 1), ("refill", 24), ("chunk", 128), ("generic_octant", 0)):
        c.set_option(k, v)


if __name__ == "__main__":
    main()
