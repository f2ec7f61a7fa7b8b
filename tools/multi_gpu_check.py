"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/multi_gpu_check.py

Every rank traces its share of the launches (launches dealt by RayTracer::ShardOwner),
the per-GPU maps are combined by uvrt_reduce (one NCCL all-reduce: sum of the photon map, max of
the max map), and the result must equal -- bit for bit -- what rank 0 gets by running all launches
alone."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sim = uv.Sim(asset_root=os.path.join(ROOT, "data"), device=local)
    sim.load_mesh("testroomopt")
    sim.init("route")
    ctx = sim.ctx
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(B.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    ctx.comm_init(bytes(idt.cpu().tolist()), rank, world)

    iters = 3
    sim.set_params(photonCount=1 << 22, maxIterations=iters)
    sim.set_shard(rank, world)
    dose = sim.run()                      # ResetDosageMap, ticks, Reduce, Shade, read-back
    pm, mx = ctx.read(uv.BUF.SUM), ctx.read(uv.BUF.MAX)
    mine = sim.rays_traced()
    t = torch.tensor([mine], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    ok = True
    if rank == 0:
        solo = uv.Sim(asset_root=os.path.join(ROOT, "data"), device=local)
        solo.load_mesh("testroomopt")
        solo.init("route")
        solo.set_params(photonCount=1 << 22, maxIterations=iters)
        dose1 = solo.run()
        c1 = solo.ctx
        pm1, mx1 = c1.read(uv.BUF.SUM), c1.read(uv.BUF.MAX)
        ok = (dose.tobytes() == dose1.tobytes() and pm.tobytes() == pm1.tobytes() and mx.tobytes() == mx1.tobytes()
              and int(t.item()) == solo.rays_traced() and sim.params.seedState == solo.params.seedState)
        print(f"[multi_gpu_check] world={world} rays total={int(t.item())} (rank0 traced {mine}) "
              f"dose/photon/max maps identical to the single-GPU run: {ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
