#!/usr/bin/env python
"""bench.py -- headline benchmark of the wavefront hot path (BASELINE.json):
Mrays/s over generate -> extend -> accumulate (+ computeDosage, dosageToColor) on the
configuration "testroomopt.glb full route.xml dose map" (configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass over the route, what one frame of the reference's MyApp::Tick computes
(myapp.cpp:156-175): for each of the 12 lamp positions of positions/route.xml one launch of
P = 2,796,202 rays (generate, extend, accumulate), then computeDosage + dosageToColor.
With N GPUs a step is N such passes (one per GPU on average: launch k of the run goes to rank
k mod N), i.e. weak scaling; the per-GPU photon maps are combined by ONE NCCL all-reduce at the end
of the run (inside the timed region).

  value     device-timed (CUDA events on the backend's stream, max over ranks), scene resident in HBM
  e2e       the same work through the reference-facing RayTracer interface with HOST buffers:
            every step uploads the scene (Tri/BVHNode/triIdx arrays -> pinned staging -> device),
            resets the maps, traces, shades and reads the dose map back; wall clock
  roofline  extend kernel: algorithmic bytes B_ray = 44 + 64*I + 52*T per ray (I, T
            measured by the oracle's traversal counters on this very workload and committed in
            profiles/traversal_stats_route.json) over the event-timed average launch duration
            (overlapping launches charged once), against the measured HBM copy bandwidth
  cpu_baseline  the reference's own kernels (oracle/_ref, compiled from its sources) or the C port,
            OpenMP over all host cores, on one pass over the route

--impl reference times only that CPU implementation (the reference has no GPU-independent build;
its OpenCL kernels compiled for the host are its CPU path -- there is no OpenCL runtime on the box).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "testroomopt.glb full route.xml dose map"
ROOM, ROUTE = "testroomopt", "route"
DATA = os.path.join(ROOT, "data")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        self.t0 = self.t1 = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.1] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's kernels on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_route_pass(room, positions, params, rays_per_launch, seed=0, want_counters=False):
    """One pass over the route on the CPU.  Returns (seconds, rays, kind, cores, counters)."""
    import uvrt_testlib as T
    tris, nodes, tri_idx, floor = room
    n = tris.shape[0]
    f32 = np.float32
    use_ref = T.ref_available() and not want_counters
    if not os.path.exists(os.path.join(T.ORACLE_DIR, "_build", "libuvrt_oracle.so")):
        T.build_checkers()
    O = T.oracle()
    R = T.ref() if use_ref else None
    pm, mx, temp = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    dose = np.zeros(n, dtype=np.float32)
    col = np.zeros((n, 9), dtype=np.float32)
    rays = np.zeros(rays_per_launch, dtype=T.RAY_DT)
    tot = T.Counters()
    t0 = time.perf_counter()
    for (x, y, dur) in positions:
        lp = (f32(x), f32(f32(floor) + f32(params.lightHeight)), f32(y))
        if use_ref:
            so = C.c_uint(0)
            R.ref_generate(T.ptr(rays), 0, rays_per_launch, lp[0], lp[1], lp[2], f32(params.lightLength), seed, C.byref(so))
            R.ref_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), rays_per_launch, n, 0)
            R.ref_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), f32(dur), n)
        else:
            so = C.c_uint32(0)
            cnt = T.Counters()
            O.orc_generate(T.ptr(rays), 0, rays_per_launch, lp[0], lp[1], lp[2], f32(params.lightLength), seed, C.byref(so))
            O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), rays_per_launch, 0, C.byref(cnt))
            O.orc_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), f32(dur), n)
            for k in ("rays", "innerVisits", "leafVisits", "triTests", "hits"):
                setattr(tot, k, getattr(tot, k) + getattr(cnt, k))
        seed = int(so.value)
    ppl = rays_per_launch
    power = f32(f32(params.lightIntensity) * f32(0.1))
    if use_ref:
        R.ref_compute_dosage(T.ptr(pm), T.ptr(dose), T.ptr(tris), ppl, power, n)
        R.ref_dosage_to_color(T.ptr(dose), T.ptr(col), f32(params.minDosage), 0, n)
    else:
        O.orc_compute_dosage(T.ptr(pm), T.ptr(dose), T.ptr(tris), ppl, power, n)
        O.orc_dosage_to_color(T.ptr(dose), T.ptr(col), f32(params.minDosage), 0, n)
    dt = time.perf_counter() - t0
    cores = int(O.orc_num_threads())
    return dt, rays_per_launch * len(positions), ("reference" if use_ref else "port"), cores, tot, seed


def traversal_stats(room, positions, params, sample_per_launch=200_000):
    """I (inner-node visits per ray) and T (triangle tests per ray) for B_ray, from the oracle's counters."""
    _, rays, _, _, cnt, _ = cpu_route_pass(room, positions, params, sample_per_launch, want_counters=True)
    return cnt.innerVisits / rays, cnt.triTests / rays


def load_room_host(uv, device=0):
    sim = uv.Sim(asset_root=DATA, device=device)
    sim.load_mesh(ROOM)
    sim.load_route(ROUTE)
    tris, nodes, tri_idx = sim.mesh_data()
    floor = sim.mesh_info()["floor"]
    return sim, (tris, nodes, tri_idx, floor)


def run_reference_arm(args):
    """The reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
    sim, room = load_room_host(uv)       # host-side loading only: no GPU, none of the CUDA path
    pos, p = sim.positions, sim.params
    # bounded sample: every step walks the whole route with 1/16 of the photons per position
    per_launch = (p.photonsPerLight // 16) & ~1
    seed = 0
    for _ in range(args.warmup):
        _, _, kind, cores, _, seed = cpu_route_pass(room, pos, p, per_launch, seed)
    t = 0.0
    rays = 0
    for _ in range(args.steps):
        dt, r, kind, cores, _, seed = cpu_route_pass(room, pos, p, per_launch, seed)
        t += dt
        rays += r
    value = rays / t / 1e6
    sample = f"{len(pos)} positions x {per_launch} rays per step (1/16 of the photons per position), {args.steps} steps"
    line = {
        "impl": "reference", "metric": "Mrays/s (extend+shade)", "value": round(value, 3), "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "room": ROOM + ".glb", "route": ROUTE + ".xml", "triangles": int(room[0].shape[0]),
                   "positions": int(len(pos)), "rays_per_launch_full": int(p.photonsPerLight),
                   "timed": "generate+extend+accumulate per position, then computeDosage+dosageToColor"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference kernels (cl/*.cl) and bvh.cpp compiled for the host by oracle/build_ref.sh; no OpenCL runtime exists on the box",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="uvrt", choices=["uvrt", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--write-traversal-stats", action="store_true", help="re-measure I and T with the oracle's counters and commit them")
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--opt", action="append", default=[], help="backend option key=value (uvrt_set_option), repeatable")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    uv = importlib.import_module("small-project-uv-robot-ray-tracer_b200")
    B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
    sim, room = load_room_host(uv, local)
    sim.init(ROUTE)                      # fails loudly without a CUDA device
    ctx = sim.ctx
    if args.variant >= 0:
        ctx.set_option("extend_variant", args.variant)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    pos, p = sim.positions, sim.params
    L, P, n_tris = len(pos), int(p.photonsPerLight), room[0].shape[0]
    rays_per_pass = L * P

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    if dist is not None:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(B.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().tolist()), rank, world)
        sim.set_shard(rank, world)

    # ---- device-timed run: scene resident, K steps, one reduction + shade + read-back at the end ----
    def run_steps(k_steps, timed):
        sim.set_params(maxIterations=k_steps * n_gpus)
        sim.reset_dosage_map()
        ctx.sync()
        total_ms = 0.0
        for s in range(k_steps):
            ctx.flush_l2()                       # cold L2 at the start of every step
            ctx.mark(0)
            for _ in range(n_gpus):              # n_gpus passes over the route, dealt out launch by launch
                sim.compute_dosage_map()         # RayTracer::ComputeDosageMap: asynchronous, like Kernel::Run
                if n_gpus == 1:
                    sim.shade()                  # one GPU: shade after every pass, as MyApp::Tick does
            if s == k_steps - 1 and n_gpus > 1:
                sim.reduce()                     # the run's one all-reduce
                sim.shade()
            ctx.mark(1)
            total_ms += ctx.elapsed_ms(0, 1)
        return total_ms

    run_steps(max(args.warmup, 3), False)
    ctx.set_option("stage_timing", 1)
    ctx.stage_time_reset()
    launches0 = ctx.launch_count()
    clocks = ClockSampler(local)
    barrier()
    clocks.begin()
    ms = run_steps(args.steps, True)
    barrier()
    clocks.end()
    clk = clocks.stop()
    launches = ctx.launch_count() - launches0
    ext_ms, ext_n = ctx.stage_time(uv.STAGE.EXTEND)
    stage_ms = {name: ctx.stage_time(getattr(uv.STAGE, name))[0] for name in
                ("GENERATE", "BIN", "EXTEND", "ACCUMULATE", "SHADE", "COLOR", "RESET")}
    ctx.set_option("stage_timing", 0)
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_rays = args.steps * n_gpus * rays_per_pass
    value = total_rays / (ms * 1e-3) / 1e6
    dose = sim.read_dose()
    traced_this_rank = sim.rays_traced()

    # ---- end to end through the RayTracer interface with host buffers ----
    tris_h, nodes_h, idx_h, _ = room
    e2e_steps = max(1, args.e2e_steps)

    def e2e_step():
        ctx.upload_scene(tris_h, nodes_h, idx_h)         # host arrays -> pinned staging -> device
        sim.set_params(maxIterations=n_gpus)
        sim.reset_dosage_map()
        for _ in range(n_gpus):
            sim.compute_dosage_map()
        sim.reduce()                                     # no-op on one GPU
        sim.shade()
        return sim.read_dose()                           # device -> host (synchronises)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        d = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = e2e_steps * n_gpus * rays_per_pass / e2e_s / 1e6
    h2d = ctx.scene_upload_bytes() + L * 12
    d2h = n_tris * 4

    # ---- whole default run (10 iterations), wall clock: the "route dose-map time" ----
    # (N GPUs: the 120 launches of the run are dealt over the ranks, one all-reduce at the end -- strong scaling)
    sim.set_params(maxIterations=10)
    sim.run()                                            # warm-up (ray buffers of this size, NCCL channels)
    barrier()
    t0 = time.perf_counter()
    sim.run()
    route_s = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([route_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        route_s = float(t.item())
    route_ms = route_s * 1e3

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (extend) ----
    peak, peak_src = load_peaks()
    # I, T of SURVEY 8(d): measured once by the oracle's traversal counters on this workload and committed
    # (profiles/traversal_stats_route.json, written by `python bench.py --write-traversal-stats`)
    spath = os.path.join(ROOT, "profiles", "traversal_stats_route.json")
    if args.write_traversal_stats or not os.path.exists(spath):
        I, T_ = traversal_stats(room, pos, p)
        if args.write_traversal_stats:
            json.dump({"workload": WORKLOAD, "inner_visits_per_ray": I, "tri_tests_per_ray": T_,
                       "source": "oracle/uvrt_oracle.c traversal counters, 12 positions x 200,000 rays of route.xml on testroomopt.glb"},
                      open(spath, "w"), indent=1)
    else:
        st = json.load(open(spath))
        I, T_ = float(st["inner_visits_per_ray"]), float(st["tri_tests_per_ray"])
    b_ray = 44.0 + 64.0 * I + 52.0 * T_
    ext_launch_ms = ext_ms / max(1, ext_n)
    achieved = b_ray * P / (ext_launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "extend_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("bytes_per_launch")
        except Exception:
            traffic = None

    line = {
        "metric": "Mrays/s (extend+shade)", "value": round(value, 1), "unit": "Mrays/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "room": ROOM + ".glb", "route": ROUTE + ".xml", "triangles": int(n_tris),
                   "positions": int(L), "rays_per_launch": int(P), "rays_per_step": int(n_gpus * rays_per_pass),
                   "timed": "generate+bin+extend+accumulate per position, computeDosage+dosageToColor (and the all-reduce) at the end of the run",
                   "l2": "flushed before every step (256 MiB memset on the same stream)",
                   "extend_variant": ctx.get_option("extend_variant"), "bin_rays": ctx.get_option("bin_rays"),
                   "pipeline": ctx.get_option("pipeline"),
                   "parallelism": f"launches dealt round-robin to {n_gpus} GPU(s), one NCCL all-reduce per run"},
        "clocks": clk,
        "e2e": {"value": round(e2e_value, 1), "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "ms_per_step": round(1e3 * e2e_s / e2e_steps, 3),
                "what": "upload_scene(host arrays) + ResetDosageMap + ComputeDosageMap + Shade + ReadDosageMap, wall clock"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "kernel": "extend", "launch_ms": round(ext_launch_ms, 4),
                     "bytes_per_ray": round(b_ray, 1), "inner_visits_per_ray": round(I, 3), "tri_tests_per_ray": round(T_, 3),
                     "peak_source": peak_src,
                     "note": "scene (6 MB) is L2/L1 resident: the HBM figure is the contract's denominator, not the binding limit (see DESIGN.md)"},
        "stage_ms_per_step": {k.lower(): round(v / args.steps, 4) for k, v in stage_ms.items()},
        "extend_mrays_s": round(P / ext_launch_ms / 1e3, 1),
        "route_dose_map_ms": round(route_ms, 2),
        "route_dose_map": "route.xml as shipped: 10 iterations x 12 positions x 2,796,202 rays, ResetDosageMap .. dose map on the host, wall clock, max over ranks",
        "dose_checksum": {"mean": float(np.mean(dose, dtype=np.float64)), "max": float(dose.max())},
    }
    if not args.no_cpu and n_gpus == 1:
        # (rank 0, N = 1 only: under torchrun the host threads are shared with the other ranks)
        # bounded CPU sample: exactly one step of the GPU workload (one pass over the route)
        per_launch = P
        dt, rays, kind, cores, _, _ = cpu_route_pass(room, pos, p, per_launch)
        line["cpu_baseline"] = {"value": round(rays / dt / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": kind,
                                "sample": f"one pass over {ROUTE}.xml, {L} positions x {per_launch} rays ({rays} rays, {dt:.1f} s)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
