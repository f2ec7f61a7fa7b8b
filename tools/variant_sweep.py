"""Times every extend variant on one launch of the default route (2,796,202 rays, lange_route
position 0 on testroomopt) with CUDA events on the context's stream.  Run on the GPU box:

    python tools/variant_sweep.py [--rays N] [--reps R] [--variants 0,1,2,...]
"""
import argparse
import importlib
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
    """Orders rays by (direction cell, origin slice): equal-probability cells in (dir.y, azimuth)."""
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
        cell = (part1by1(t) | (part1by1(ph) << 1)).astype(np.int64)
        key = cell * nY + y
    elif mode == "ytp":
        key = (y * nT + t) * nP + ph
    elif mode == "ytps":      # serpentine azimuth: neighbouring bins stay neighbours across rows
        phs = np.where(t % 2 == 0, ph, nP - 1 - ph)
        key = (y * nT + t) * nP + phs
    elif mode == "morton3":   # interleave y, t, ph (equal bit counts expected)
        def p3(x):
            x = x.astype(np.uint64) & 0x3ff
            x = (x | (x << 16)) & 0x30000ff
            x = (x | (x << 8)) & 0x300f00f
            x = (x | (x << 4)) & 0x30c30c3
            x = (x | (x << 2)) & 0x9249249
            return x
        key = (p3(t) | (p3(ph) << 1) | (p3(y) << 2)).astype(np.int64)
    elif mode == "tyq":       # direction cell major, then origin slice (fine), then azimuth
        key = (t * nY + y) * nP + ph
    else:
        key = (t * nP + ph) * nY + y
    return gen[np.argsort(key, kind="stable")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=2796202)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variants", default="0,1,2,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24")
    ap.add_argument("--hist", default="0,1")
    ap.add_argument("--bps", default="0")
    ap.add_argument("--refill", default="24")
    ap.add_argument("--cfg", default="0", help="simple_cfg values (block size / register cap of the simple kernel)")
    ap.add_argument("--positions", default="0")
    ap.add_argument("--bin", default="0", help="device binning configs: 0 (off) or nY,nT,nP; ';'-separated")
    ap.add_argument("--sort", default="none", help="host-side ray orderings to try: none or nT,nP,nY[,morton]; ';'-separated")
    args = ap.parse_args()
    sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
    sim.load_mesh("testroomopt")
    sim.init("lange_route")
    c = sim.ctx
    floor = sim.mesh_info()["floor"]
    p = sim.params
    pos = sim.positions
    print(json.dumps({"device": c.device_info(), "scene": c.scene_info()}))
    P = args.rays
    for pi in [int(x) for x in args.positions.split(",")]:
        lp = (np.float32(pos[pi, 0]), np.float32(np.float32(floor) + np.float32(p.lightHeight)), np.float32(pos[pi, 1]))
        c.generate(lp, p.lightLength, 0, P, 0)
        gen = c.read(uv.BUF.RAYS, P)
        ref_counts = None
        gen0 = gen
        for sort in args.sort.split(";"):
          gen = sorted_rays(gen0, sort, lp, p.lightLength)
          for binspec in args.bin.split(";"):
           if binspec == "0":
               c.set_option("bin_rays", 0)
           else:
               by, bt, bp = [int(x) for x in binspec.split(",")]
               c.set_option("bin_rays", 1); c.set_option("bin_y", by); c.set_option("bin_t", bt); c.set_option("bin_p", bp)
           for v in [int(x) for x in args.variants.split(",")]:
               for hist in [int(x) for x in args.hist.split(",")]:
                   if hist and v < 10:
                       continue
                   for bps, refill in [(int(x), int(y)) for x in args.bps.split(",") for y in (args.refill.split(",") if v >= 10 else args.cfg.split(","))]:
                       c.set_option("refill" if v >= 10 else "simple_cfg", refill)
                       c.set_option("extend_variant", v)
                       c.set_option("hist_mode", hist)
                       c.set_option("blocks_per_sm", bps)
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
                       ok = bool(np.array_equal(counts, ref_counts))
                       best = min(times)
                       print(json.dumps({"pos": pi, "sort": sort, "bin": binspec, "variant": v, "hist": hist, "bps": bps, "refill_or_cfg": refill, "ms_best": round(best, 4),
                                         "ms_med": round(float(np.median(times)), 4), "mrays_s": round(P / best / 1e3, 1),
                                         "counts_equal": ok}), flush=True)
    c.set_option("extend_variant", -1)
    c.set_option("bin_rays", 1)


if __name__ == "__main__":
    main()
