"""Where does the wall clock of the route dose-map run go?  (apportion the fixed cost at 8 GPUs)

    python [-m torch.distributed.run --nproc-per-node N ...] tools/timeline_run.py [--parts K] [--out gpurun_out/timeline]

Runs the default run of route.xml (335,544,240 rays) shared between the ranks with the backend's "timeline" option
on: every C-ABI call (host clock) and every stage launch (host time of the enqueue, device start/stop).  Each rank
writes <out>_n<N>_rank<r>.json; rank 0 prints a summary: wall clock, device busy time of extend, idle gaps, time
before the first kernel and after the last one, time inside the all-reduce + fold, per-call host cost.
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
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")


def summarise(path, wall_ms):
    t = json.load(open(path))
    k = t["kernels"]
    calls = t["calls"]
    ext = [(a, z) for (name, h, a, z) in k if name == "extend"]
    allk = sorted([(a, z) for (_, _, a, z) in k])
    busy, cur_a, cur_z = 0.0, None, None
    for a, z in allk:                      # union of the device intervals
        if cur_z is None or a > cur_z:
            if cur_z is not None:
                busy += cur_z - cur_a
            cur_a, cur_z = a, z
        else:
            cur_z = max(cur_z, z)
    if cur_z is not None:
        busy += cur_z - cur_a
    per_call = {}
    for name, a, z in calls:
        c = per_call.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += z - a
    first_call = min(a for _, a, _ in calls) if calls else 0.0
    last_call = max(z for _, _, z in calls) if calls else 0.0
    return {"wall_ms": round(wall_ms, 3), "kernels": len(k), "device_first_start_us": round(allk[0][0], 1) if allk else None,
            "device_last_stop_us": round(allk[-1][1], 1) if allk else None, "device_busy_us": round(busy, 1),
            "extend_busy_us": round(sum(z - a for a, z in ext), 1), "extend_launches": len(ext),
            "host_first_call_us": round(first_call, 1), "host_last_call_end_us": round(last_call, 1),
            "host_us_per_call": {n: [c[0], round(c[1], 1)] for n, c in sorted(per_call.items())}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", type=int, default=0)
    ap.add_argument("--cost-aware", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline"))
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sim = uv.Sim(asset_root=os.path.join(ROOT, "data"), device=local)
    sim.load_mesh("testroomopt")
    sim.init("route")
    ctx = sim.ctx
    if dist is not None:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(B.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().tolist()), rank, world)
        sim.set_shard(rank, world)
        sim.set_shard_parts(args.parts)
        sim.set_cost_aware(bool(args.cost_aware))
    for _ in range(2):
        sim.run()
    best = None
    for rep in range(3):
        sim.set_seed(0)
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        ctx.set_option("timeline", 1)
        t0 = time.perf_counter()
        sim.run()
        wall = (time.perf_counter() - t0) * 1e3
        path = f"{args.out}_n{world}_rank{rank}.json"
        if best is None or wall < best:
            best = wall
            ctx.timeline_dump(path)
        ctx.set_option("timeline", 0)
    s = summarise(f"{args.out}_n{world}_rank{rank}.json", best)
    s.update({"n_gpus": world, "rank": rank, "parts": sim.shard_parts(), "cost_aware": int(bool(args.cost_aware) and args.parts == 0)})
    if dist is not None:
        import torch
        gathered = [None] * world
        dist.all_gather_object(gathered, s)
        if rank == 0:
            for g in gathered:
                print(json.dumps(g), flush=True)
        dist.destroy_process_group()
    else:
        print(json.dumps(s), flush=True)
    sim.close()


if __name__ == "__main__":
    main()
