"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (tests/test_gpu_multi.py does):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29533 tools/multi_gpu_check.py

1. The default run of BASELINE configs[1] (route.xml, 10 iterations, 335,544,240 rays) shared between the ranks -- the
   cost-aware plan of whole launches (default), the rotation with its automatic choice of ranges, whole launches
   (parts 1) and launches cut into ray ranges (parts 2 and 4): photon map, max map, dose and
   colours must hash to the golden the reference's own compiled sources produced (tests/golden/route_runs.json).
2. A smaller run with awkward durations: the sharded result must equal, bit for bit, what rank 0 gets alone.
Every rank traces its units (RayTracer::ShardOwner), the integer count rows are summed by ONE ncclAllReduce per
window and folded in launch order (uvrt_matrix_fold)."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")


def fnv(a):
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).view(np.uint8).reshape(-1).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xffffffffffffffff
    return f"{h:016x}"


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
    sim.set_shard(rank, world)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "route_runs.json")))["runs"]["route"]["after_iteration"][9]
    ok = True
    for parts in (0, -1, 1, 2, 4):         # 0: cost-aware plan of whole launches (default); -1: rotation + AutoParts; else fixed ranges
        sim.set_cost_aware(parts == 0)
        sim.set_shard_parts(max(parts, 0))
        sim.set_seed(0)
        dose = sim.run()                      # ResetDosageMap, all passes, Reduce (all-reduce + fold), Shade, read-back
        got = (fnv(ctx.read(uv.BUF.SUM)), fnv(ctx.read(uv.BUF.MAX)), fnv(dose), fnv(ctx.read(uv.BUF.COLOR)), int(sim.params.seedState))
        want = (golden["fnv_photonMap"], golden["fnv_maxPhotonMap"], golden["fnv_dose"], golden["fnv_color"], golden["seed"])
        same = got == want
        ok = ok and same
        t = torch.tensor([sim.rays_traced()], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        ok = ok and int(t.item()) == 335_544_240
        if rank == 0:
            print(f"[multi_gpu_check] world={world} parts={parts} (in effect {sim.shard_parts()}): default run, dose fnv {got[2]} "
                  f"{'== golden' if same else '!= golden ' + want[2]}; rays over all ranks {int(t.item())}", flush=True)
    # a smaller run with durations whose products do not add exactly
    pos = sim.positions.copy()
    pos[:, 2] = np.array([0.1, 7.3, 60.0, 1e-3, 33.3, 2.5, 0.7, 19.0, 5.5, 0.01, 100.0, 1.0], dtype=np.float32)
    sim.set_positions(pos)
    sim.set_params(photonCount=1 << 23, maxIterations=3)
    sim.set_shard_parts(3)
    sim.set_seed(12345)
    dose = sim.run()
    pm, mx = ctx.read(uv.BUF.SUM), ctx.read(uv.BUF.MAX)
    if rank == 0:
        solo = uv.Sim(asset_root=os.path.join(ROOT, "data"), device=local)
        solo.load_mesh("testroomopt")
        solo.init("route")
        solo.set_positions(pos)
        solo.set_params(photonCount=1 << 23, maxIterations=3)
        solo.set_seed(12345)
        dose1 = solo.run()
        c1 = solo.ctx
        same = (dose.tobytes() == dose1.tobytes() and pm.tobytes() == c1.read(uv.BUF.SUM).tobytes()
                and mx.tobytes() == c1.read(uv.BUF.MAX).tobytes() and sim.params.seedState == solo.params.seedState)
        ok = ok and same
        print(f"[multi_gpu_check] world={world} parts=3, awkward durations: maps identical to the single-GPU run: {same}", flush=True)
        solo.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(flag.item()) else "MULTI_GPU_MISMATCH", flush=True)
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
