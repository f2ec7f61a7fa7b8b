"""The N > 1 contract on the CPU (gloo, world_size 2): launches are dealt round-robin, rotated once
per pass (RayTracer::ShardOwner, called here through libuvrt_host), every rank advances the SEED
chain for all launches, and ONE reduction -- sum of the photon maps, max of the per-launch maxima --
reproduces the single-rank result exactly.  The per-launch work is done by the oracle here; the same
check runs on real GPUs over NCCL in tools/multi_gpu_check.py."""
import os
import subprocess
import sys

import numpy as np

import uvrt_testlib as T

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["UVRT_TESTS"])
import uvrt_testlib as T
import importlib
sys.path.insert(0, T.ROOT)
H = importlib.import_module("small-project-uv-robot-ray-tracer_b200").host()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
d = np.load(os.environ["UVRT_SCENE"])
tris, nodes, tri_idx = d["tris"], d["nodes"].view(T.NODE_DT).reshape(-1), d["tri_idx"]
pos, floor = d["pos"], np.float32(d["floor"])
O = T.oracle()
n = tris.shape[0]; P = 40000; f32 = np.float32
def run(world, rank):
    pm, mx, temp = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    rays = np.zeros(P, dtype=T.RAY_DT); seed = 0; k = 0
    for it in range(2):
        for (x, y, dur) in pos:
            so = C.c_uint32(0)
            lp = (f32(x), f32(floor + f32(0.6)), f32(y))
            if H.uvrt_host_shard_owner(k, len(pos), world) == rank:
                O.orc_generate(T.ptr(rays), 0, P, lp[0], lp[1], lp[2], f32(1.0), seed, C.byref(so))
                O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, 1, None)
                O.orc_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), f32(dur), n)
            else:
                O.orc_generate(T.ptr(rays), 0, 0, lp[0], lp[1], lp[2], f32(1.0), seed, C.byref(so))   # SEED only
            seed = int(so.value); k += 1
    return pm, mx, seed
pm, mx, seed = run(world, rank)
tp, tm = torch.from_numpy(pm), torch.from_numpy(mx)
dist.all_reduce(tp, op=dist.ReduceOp.SUM)
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    pm1, mx1, seed1 = run(1, 0)
    ok = pm1.tobytes() == tp.numpy().tobytes() and mx1.tobytes() == tm.numpy().tobytes() and seed == seed1
    print("MULTIRANK_OK" if ok else "MULTIRANK_MISMATCH", flush=True)
dist.destroy_process_group()
'''


def test_round_robin_shards_reduce_to_single_rank_result(room, tmp_path, checkers):
    tris, nodes, tri_idx, floor = room
    pos = np.array([[-0.255, -3.31, 60.0], [0.085, -2.46, 30.0], [-0.51, -1.19, 0.1]], dtype=np.float32)
    scene = str(tmp_path / "scene.npz")
    np.savez(scene, tris=tris, nodes=nodes.view(np.uint8), tri_idx=tri_idx, pos=pos, floor=np.float32(floor))
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, UVRT_TESTS=os.path.dirname(os.path.abspath(__file__)), UVRT_SCENE=scene, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(worker)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MULTIRANK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_shard_owner_rotates_positions_over_ranks():
    """Every position must visit every rank (L = 12, N = 8 with plain k mod N would pin three positions
    to each rank), every launch has exactly one owner, and ranks get equal shares."""
    import importlib
    H = importlib.import_module("small-project-uv-robot-ray-tracer_b200").host()
    for L, N in ((12, 8), (12, 4), (12, 2), (12, 3), (5, 5), (7, 8), (1, 4), (12, 1)):
        passes = 4 * N
        seen = {}
        share = [0] * N
        for k in range(L * passes):
            r = H.uvrt_host_shard_owner(k, L, N)
            assert 0 <= r < N
            seen.setdefault(k % L, set()).add(r)
            share[r] += 1
        assert all(len(v) == N for v in seen.values()), (L, N, seen)
        assert max(share) - min(share) <= 1 + L * passes // (N * 50), (L, N, share)
