"""How much does cutting a launch into ray ranges cost?  (the policy behind RayTracer::AutoParts)

    python tools/split_sweep.py [--parts 1,2,4,8] [--bins default|scaled]

One pass over route.xml (12 launches of 2,796,202 rays) traced through uvrt_trace_row as 12 * parts launches of
P / parts rays each -- what ONE rank would do if it had to trace everything in pieces of that size -- device-timed
(events on the context's stream), best of 3.  The count rows must be identical for every split.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", default="1,2,4,8,16")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    sim = uv.Sim(asset_root=os.path.join(ROOT, "data"))
    sim.load_mesh("testroomopt")
    sim.init("route")
    ctx = sim.ctx
    pos, p = sim.positions, sim.params
    floor = sim.mesh_info()["floor"]
    f32 = np.float32
    P = int(p.photonsPerLight)
    lamps = [(f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y)) for x, y, _ in pos]
    want = None
    for scaled in (0, 1):
        for parts in [int(x) for x in args.parts.split(",")]:
            if scaled and parts == 1:
                continue
            if scaled:      # keep ~43 rays per bin: fewer azimuth / elevation cells for smaller launches
                shrink = parts
                bt, bp = 32, 128
                while shrink > 1 and bp > 16:
                    bp //= 2; shrink //= 2
                    if shrink > 1:
                        bt //= 2; shrink //= 2
                ctx.set_option("bin_t", max(bt, 4)); ctx.set_option("bin_p", bp)
            else:
                ctx.set_option("bin_t", 32); ctx.set_option("bin_p", 128)
            best = 1e9
            for rep in range(args.reps + 1):
                ctx.matrix_begin(len(lamps))
                ctx.sync()
                ctx.flush_l2()
                ctx.mark(0)
                for row, lp in enumerate(lamps):
                    for j in range(parts):
                        first, last = P * j // parts, P * (j + 1) // parts
                        ctx.trace_row(row, lp, p.lightLength, first, last - first, 17 * row)
                ctx.mark(1)
                ms = ctx.elapsed_ms(0, 1)
                if rep:
                    best = min(best, ms)
            m = ctx.read(uv.BUF.MATRIX, len(lamps))
            if want is None:
                want = m
            print(json.dumps({"parts": parts, "rays_per_launch": P // parts, "bins": ctx.get_option("bin_y") * ctx.get_option("bin_t") * ctx.get_option("bin_p"),
                              "ms_per_pass": round(best, 4), "mrays_s": round(len(lamps) * P / best / 1e3, 1),
                              "rows_equal": bool(np.array_equal(m, want))}), flush=True)
    sim.close()


if __name__ == "__main__":
    main()
