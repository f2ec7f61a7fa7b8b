"""BASELINE.json config 5: synthetic 10 M-triangle soup (incoherent traversal stress), ray sweep 1e6 .. 1e9 at
1/2/4/8 GPUs.  Run alone (1 GPU) or under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/soup_sweep.py [--tris 10000000] [--rays 1000000,10000000,100000000,1000000000]

Every rank builds the BVH on its GPU (uvrt_build_bvh), uploads the scene, and traces its share of the rays:
a sweep point of R rays is cut into launches of at most 2^23 rays (above 15.79 M rays per
launch the reference's float-rounded seed expression repeats rays); launch j goes to rank j mod N, lamp
position j mod 12 of the soup route, SEED j.  The per-GPU integer counts are summed with one
uvrt_reduce_counts (NCCL) per sweep point.  Time = CUDA events on the context's stream around the launches
and the reduction, max over ranks.  Counts are integers, so the N-GPU result must equal the 1-GPU result
exactly: rank 0 prints an FNV hash of the reduced counts to compare across N.
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
B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
from soup import make_soup, soup_route  # noqa: E402


def fnv(a):
    h = 1469598103934665603
    # hash of 64-bit partial sums keeps this cheap for 10 M counters
    for v in np.add.reduceat(a.astype(np.uint64) * (np.arange(a.size, dtype=np.uint64) % 1021 + 1), np.arange(0, a.size, 4096)):
        h = ((h ^ int(v)) * 1099511628211) & 0xffffffffffffffff
    return f"{h:016x}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--rays", default="1000000,10000000,100000000,1000000000")
    ap.add_argument("--launch", type=int, default=1 << 23)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = uv.Context(local)
    if dist is not None:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(B.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().tolist()), rank, world)
    t0 = time.perf_counter()
    soup = make_soup(args.tris)
    t1 = time.perf_counter()
    tris, nodes, tri_idx = ctx.build_bvh(soup)
    t2 = time.perf_counter()
    ctx.upload_scene(tris, nodes, tri_idx)
    ctx.sync()
    t3 = time.perf_counter()
    route = soup_route()
    if rank == 0:
        print(json.dumps({"triangles": args.tris, "n_gpus": world, "scene": ctx.scene_info(), "make_s": round(t1 - t0, 2),
                          "device_bvh_build_s": round(t2 - t1, 3), "upload_s": round(t3 - t2, 3)}), flush=True)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for R in [int(x) for x in args.rays.split(",")]:
        launches = [(j, min(args.launch, R - j * args.launch)) for j in range((R + args.launch - 1) // args.launch)]
        mine = [(j, n) for (j, n) in launches if j % world == rank]
        for rep in range(2):                      # the first repetition warms up (ray buffers, clocks)
            ctx.reset(False)
            barrier()
            ctx.mark(0)
            for j, n in mine:
                x, z, _ = route[j % len(route)]
                ctx.trace_counts((np.float32(x), np.float32(0.5), np.float32(z)), 1.0, 0, n, j)
            if world > 1:
                ctx.reduce_counts()
            ctx.mark(1)
            ms = ctx.elapsed_ms(0, 1)
            barrier()
        if dist is not None:
            import torch
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rank == 0:
            counts = ctx.read(uv.BUF.COUNTS)
            print(json.dumps({"rays": R, "n_gpus": world, "launches": len(launches), "ms": round(ms, 3),
                              "mrays_s": round(R / ms / 1e3, 1), "hits": int(counts.astype(np.int64).sum()),
                              "counts_fnv": fnv(counts)}), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
