"""Parity of the CUDA path (through the C ABI of include/uvrt.h / uvrt_host.h) against the oracle
on identical inputs.  Everything here needs a B200: run with `-m gpu`.

Bars: generate 32-byte records bit-equal; extend (dist bits, triID)
equal for EVERY ray and per-triangle counts integer-equal; accumulate / computeDosage /
dosageToColor / reset bit-equal (stricter than the 1-ulp allowance); end-to-end dose within
rel. 1e-3 (north_star) -- and in fact bit-equal, which is asserted too."""
import ctypes as C
import os

import numpy as np
import pytest

import uvrt_testlib as T
from conftest import lange_pos0

pytestmark = pytest.mark.gpu

DOSE_RTOL = 1e-3          # tolerance stated by BASELINE.json north_star
SIMPLE_VARIANTS = [0, 1, 2, 50]      # 50: the certified fast extend (csrc/uvrt_fast.cuh)
PERSIST_VARIANTS = [10, 11, 12, 16, 17, 19, 20, 22, 23, 24]


@pytest.fixture(scope="module")
def ctx(uv, room):
    c = uv.Context(0)
    tris, nodes, tri_idx, _ = room
    c.upload_scene(tris, nodes, tri_idx)
    yield c
    c.close()


def oracle_launch(room, lp, P, seed_in=0, first=0, light_length=1.0):
    tris, nodes, tri_idx, _ = room
    O = T.oracle()
    rays = np.zeros(P, dtype=T.RAY_DT)
    so = C.c_uint32(0)
    O.orc_generate(T.ptr(rays), first, P, lp[0], lp[1], lp[2], light_length, seed_in, C.byref(so))
    gen = rays.copy()
    temp = np.zeros(tris.shape[0], dtype=np.int32)
    cnt = T.Counters()
    O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, 0, C.byref(cnt))
    return gen, rays, temp, int(so.value), cnt


def test_device_is_blackwell(uv, ctx):
    info = ctx.device_info()
    assert info["cc"][0] >= 10, info
    s = ctx.scene_info()
    assert s["inner"] + s["leaves"] == 89417 and s["leaves"] == 44709 and s["depth"] == 21


def test_generate_bit_exact_and_golden(uv, ctx, room, golden):
    lp = lange_pos0(room[3])
    P = 1000000
    ctx.generate(lp, 1.0, 0, P, 0)
    got = ctx.read(uv.BUF.RAYS, P)
    want, _, _, seed_out, _ = oracle_launch(room, lp, P)
    assert got.tobytes() == want.tobytes()
    assert f"{T.fnv(got):016x}" == golden["launch"][str(P)]["fnv_rays_after_generate"]
    chain = ctx.seed_chain([lp], 1.0, 0)
    assert int(chain[1]) == seed_out == golden["launch"][str(P)]["seed_out"]


@pytest.mark.parametrize("first,seed_in,length", [(0, 0xdeadbeef, 1.0), (5_000_001, 435838349, 0.37), (126_000_000, 1, 2.5)])
def test_generate_shards_reproduce_global_ids(uv, ctx, room, first, seed_in, length):
    lp = (np.float32(0.51000142), np.float32(-0.7), np.float32(-0.25500044))
    P = 300_001
    ctx.generate(lp, length, first, P, seed_in)
    got = ctx.read(uv.BUF.RAYS, P)
    want = oracle_launch(room, lp, P, seed_in, first, length)[0]
    assert got.tobytes() == want.tobytes()


def test_seed_chain_matches_golden(uv, ctx, room, golden):
    g = golden["pass_lange_route"]
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_route("lange_route")
    pos, p = sim.positions, sim.params
    f32 = np.float32
    lps = [(f32(x), f32(f32(room[3]) + f32(p.lightHeight)), f32(y)) for x, y, _ in pos]
    chain = ctx.seed_chain(lps, p.lightLength, 0)
    assert [int(s) for s in chain] == g["seed_chain"]


@pytest.mark.parametrize("variant", SIMPLE_VARIANTS + PERSIST_VARIANTS)
@pytest.mark.parametrize("hist,binned", [(0, 1), (1, 1), (0, 0)])
def test_extend_bit_exact(uv, ctx, room, golden, variant, hist, binned):
    if hist and variant < 10:
        pytest.skip("hist_mode only exists for the persistent kernels")
    if variant >= 10 and not ctx.get_option("experiments"):
        pytest.skip("rejected variants are only compiled with `make EXPERIMENTS=1`")
    lp = lange_pos0(room[3])
    P = 1000000
    gen, want, want_counts, _, _ = oracle_launch(room, lp, P)
    ctx.set_option("extend_variant", variant)
    ctx.set_option("hist_mode", hist)
    ctx.set_option("bin_rays", binned)
    try:
        ctx.reset(True)
        ctx.write(uv.BUF.RAYS, gen)
        ctx.extend(P)
        got = ctx.read(uv.BUF.RAYS, P)
        counts = ctx.read(uv.BUF.COUNTS)
    finally:
        ctx.set_option("extend_variant", -1)
        ctx.set_option("hist_mode", 0)
        ctx.set_option("bin_rays", 1)
    bad = np.flatnonzero((got["dist"].view(np.uint32) != want["dist"].view(np.uint32)) | (got["triID"] != want["triID"]))
    assert bad.size == 0, f"{bad.size} rays differ, first {bad[:5]}: {got[bad[:3]]} vs {want[bad[:3]]}"
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(counts, want_counts)
    g = golden["launch"][str(P)]
    assert f"{T.fnv(counts):016x}" == g["fnv_counts"]
    assert f"{int(T.oracle().orc_fnv_hits(T.ptr(got), P)):016x}" == g["fnv_hits"]


def test_full_size_launch_matches_golden(uv, ctx, room, golden):
    """The launch size of the default route: 2,796,202 rays, default kernel, fused trace call."""
    lp = lange_pos0(room[3])
    P = 2796202
    g = golden["launch"][str(P)]
    ctx.reset(True)
    ctx.trace_counts(lp, 1.0, 0, P, 0)
    rays = ctx.read(uv.BUF.RAYS, P)
    counts = ctx.read(uv.BUF.COUNTS)
    assert f"{int(T.oracle().orc_fnv_hits(T.ptr(rays), P)):016x}" == g["fnv_hits"]
    assert f"{T.fnv(counts):016x}" == g["fnv_counts"]
    assert int(counts.sum()) == g["hits"] == int((rays["dist"] != np.float32(1e30)).sum())
    assert [int(counts.argmax()), int(counts.max())] == g["hottest"]
    for i, hexs in g["samples"].items():
        assert rays[int(i)].tobytes().hex() == hexs


@pytest.mark.parametrize("variant", [0, 1, 2, 50, 11, 23])
@pytest.mark.parametrize("binned", [0, 1])
def test_extend_degenerate_rays(uv, ctx, room, variant, binned):
    """Axis-parallel directions (division by zero, 0/0 = NaN on slab planes), origins outside the
    room, zero-length directions: the strict path must reproduce the oracle's IEEE behaviour."""
    if variant >= 10 and not ctx.get_option("experiments"):
        pytest.skip("rejected variants are only compiled with `make EXPERIMENTS=1`")
    tris, nodes, tri_idx, floor = room
    rng = np.random.default_rng(5)
    n = 70000 if binned else 20000      # binning starts at 65,536 rays
    rays = np.zeros(n, dtype=T.RAY_DT)
    rays["orig"] = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:6000, 0] = 0.0                                # one exact zero
    d[2000:8000, 2] = 0.0                            # two exact zeros for 2000..5999
    d[8000:8100] = 0.0                               # null directions
    d[8050:8100, 1] = np.nan                          # NaN directions
    d[8100:8200] *= 1e-20                            # tiny directions (not 'tame')
    rays["dir"] = d.astype(np.float32)
    boxmin = nodes[0]["min"]
    rays["orig"][:3000, 0] = boxmin[0]               # on a slab plane with dir.x == 0 -> 0/0
    rays["orig"][9000:9500] = 0.0                    # exact zero origin
    rays["orig"][9500:10000] *= 1e-9                 # tiny origin (not 'tame')
    rays["orig"][10000:10500] += 50.0                # outside the room
    rays["dist"] = 1e30
    want = rays.copy()
    wc = np.zeros(tris.shape[0], dtype=np.int32)
    T.oracle().orc_extend(T.ptr(wc), T.ptr(tris), T.ptr(want), T.ptr(nodes), T.ptr(tri_idx), n, 0, None)
    ctx.set_option("extend_variant", variant)
    ctx.set_option("bin_rays", binned)
    try:
        ctx.reset(True)
        ctx.write(uv.BUF.RAYS, rays)
        ctx.extend(n)
        got = ctx.read(uv.BUF.RAYS, n)
        counts = ctx.read(uv.BUF.COUNTS)
    finally:
        ctx.set_option("extend_variant", -1)
        ctx.set_option("bin_rays", 1)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(counts, wc)


def test_per_triangle_passes_bit_exact(uv, ctx, room):
    tris = room[0]
    n = tris.shape[0]
    O = T.oracle()
    rng = np.random.default_rng(9)
    pm, mx = np.zeros(n), np.zeros(n)
    ctx.reset(True)
    for dur in (60.0, 0.1, 7.25):
        temp = rng.integers(0, 70000, n).astype(np.int32)
        temp[::5] = 0
        ctx.write(uv.BUF.COUNTS, temp)
        ctx.accumulate(dur)
        O.orc_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), np.float32(dur), n)
        assert not ctx.read(uv.BUF.COUNTS).any()
    assert ctx.read(uv.BUF.SUM).tobytes() == pm.tobytes()
    assert ctx.read(uv.BUF.MAX).tobytes() == mx.tobytes()
    for use_max, ppl, power, minv, thr in ((0, 2796202, 44.019705, 100.0, 0), (1, 2796202, 44019.705, 1500.0, 1), (0, 1, 1.0, 1e-3, 1)):
        ctx.shade(use_max, ppl, power)
        ctx.color(minv, thr)
        dose = np.zeros(n, dtype=np.float32)
        O.orc_compute_dosage(T.ptr(mx if use_max else pm), T.ptr(dose), T.ptr(tris), ppl, np.float32(power), n)
        col = np.zeros((n, 9), dtype=np.float32)
        O.orc_dosage_to_color(T.ptr(dose), T.ptr(col), np.float32(minv), thr, n)
        assert ctx.read(uv.BUF.DOSE).tobytes() == dose.tobytes()
        assert ctx.read(uv.BUF.COLOR).tobytes() == col.tobytes()
    ctx.reset(False)
    assert not ctx.read(uv.BUF.SUM).any() and not ctx.read(uv.BUF.MAX).any()
    assert ctx.read(uv.BUF.COLOR).any()
    ctx.reset(True)
    assert not ctx.read(uv.BUF.COLOR).any()


def oracle_route_run(room, pos, p, iterations, photons_per_light, seed=0):
    tris, nodes, tri_idx, floor = room
    O = T.oracle()
    n = tris.shape[0]
    f32 = np.float32
    pm, mx, temp = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    rays = np.zeros(photons_per_light, dtype=T.RAY_DT)
    total = 0
    for _ in range(iterations):
        for (x, y, dur) in pos:
            so = C.c_uint32(0)
            O.orc_generate(T.ptr(rays), 0, photons_per_light, f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y),
                           f32(p.lightLength), seed, C.byref(so))
            O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), photons_per_light, 0, None)
            O.orc_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), f32(dur), n)
            seed = int(so.value)
            total += photons_per_light
    dose = np.zeros(n, dtype=np.float32)
    O.orc_compute_dosage(T.ptr(pm), T.ptr(dose), T.ptr(tris), total // len(pos), f32(f32(p.lightIntensity) * f32(0.1)), n)
    col = np.zeros((n, 9), dtype=np.float32)
    O.orc_dosage_to_color(T.ptr(dose), T.ptr(col), f32(p.minDosage), 0, n)
    return pm, mx, dose, col, seed


@pytest.mark.parametrize("route", ["route", "lange_route"])
def test_raytracer_end_to_end(uv, room, route):
    """The drop-in RayTracer (LoadMesh -> Init -> ResetDosageMap -> Tick loop) against the oracle."""
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    sim.init(route)
    sim.set_params(photonCount=1 << 22, maxIterations=2)
    p = sim.params
    assert p.photonsPerLight == ((1 << 22) // 12) & ~1
    dose = sim.run()
    c = sim.ctx
    pm, mx, want, col, seed = oracle_route_run(room, sim.positions, p, 2, p.photonsPerLight)
    nz = want != 0
    assert np.array_equal(nz, dose != 0)
    rel = np.abs(dose[nz].astype(np.float64) - want[nz]) / want[nz]
    assert rel.max() <= DOSE_RTOL
    assert dose.tobytes() == want.tobytes()          # in fact bit-identical
    assert c.read(uv.BUF.SUM).tobytes() == pm.tobytes()
    assert c.read(uv.BUF.MAX).tobytes() == mx.tobytes()
    assert c.read(uv.BUF.COLOR).tobytes() == col.tobytes()
    q = sim.params
    assert q.seedState == seed and q.currIterations == 2 and q.finishedComputation
    assert q.photonMapSize == 2 * 12 * p.photonsPerLight == sim.rays_traced()
    sim.close()


def test_full_pass_matches_golden(uv, room, golden):
    """One pass over lange_route at the default size (12 x 2,796,202 rays) against the vectors
    produced by the reference's own compiled sources (App. C.2)."""
    g = golden["pass_lange_route"]
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    sim.init("lange_route")
    sim.set_params(maxIterations=1)
    dose = sim.run()
    c = sim.ctx
    assert f"{T.fnv(c.read(uv.BUF.SUM)):016x}" == g["fnv_photonMap"]
    assert f"{T.fnv(c.read(uv.BUF.MAX)):016x}" == g["fnv_maxPhotonMap"]
    assert f"{T.fnv(dose):016x}" == g["fnv_dose"]
    assert f"{T.fnv(c.read(uv.BUF.COLOR)):016x}" == g["fnv_color"]
    assert [int(v) for v in dose[:8].view(np.uint32)] == g["dose_head_bits"]
    assert int((dose == 0).sum()) == g["unlit"]
    assert sim.params.seedState == g["seed_chain"][-1]
    # energy bookkeeping at full size: every hit landed in exactly one triangle's sum
    assert int(round(c.read(uv.BUF.SUM).sum() / 60.0)) == sum(g["hits_per_position"])
    sim.close()


@pytest.mark.parametrize("deal", ["rotation", "cost_aware"])
def test_sharded_run_equals_single(uv, room, deal):
    """Whole launches shared between two 'ranks' (here: two contexts on one GPU whose maps are combined on the host) --
    by the rotation of RayTracer::ShardOwner, or by the cost-aware plan (every rank probes the lamp positions on its own
    and must arrive at the same deal) -- reproduce the single-context maps exactly."""
    results = []
    for rank, count in ((0, 1), (0, 2), (1, 2)):
        sim = uv.Sim(asset_root=T.DATA)
        sim.load_mesh("testroomopt")
        sim.init("route")
        sim.set_params(photonCount=1 << 21, maxIterations=3)
        sim.set_shard(rank, count)
        sim.set_shard_parts(1 if deal == "rotation" else 0)   # whole launches either way: per-rank maxima combine with max()
        sim.set_cost_aware(deal == "cost_aware")
        sim.reset_dosage_map()
        while not sim.tick():
            pass
        sim.reduce()                             # no communicator: folds this "rank's" rows of the count matrix
        c = sim.ctx
        results.append((c.read(uv.BUF.SUM), c.read(uv.BUF.MAX), sim.params.seedState, sim.params.photonMapSize, sim.rays_traced()))
        sim.close()
        sim = None
    one, a, b = results
    assert np.array_equal(one[0], a[0] + b[0])
    assert np.array_equal(one[1], np.maximum(a[1], b[1]))
    assert one[2] == a[2] == b[2] and one[3] == a[3] == b[3]
    assert one[4] == a[4] + b[4]


def test_split_launch_counts_add_up(uv, ctx, room):
    """A launch cut into ray ranges (firstRay) gives the same counts as the whole launch."""
    lp = lange_pos0(room[3])
    P = 500_000
    ctx.reset(True)
    ctx.trace_counts(lp, 1.0, 0, P, 77)
    whole = ctx.read(uv.BUF.COUNTS)
    ctx.reset(True)
    parts = np.zeros_like(whole)
    for first, n in ((0, 123_457), (123_457, 0), (123_457, 300_000), (423_457, 76_543)):
        ctx.trace_counts(lp, 1.0, first, n, 77)
    parts = ctx.read(uv.BUF.COUNTS)
    assert np.array_equal(whole, parts)


def test_calibration_scene_and_analytic_irradiance(uv, room):
    """CalibratePower (raytracer.cpp:151-227): a 2-triangle scene whose BVH root is a leaf.  With
    power 1 the simulated irradiance on the patch must match the closed form for a uniformly
    radiating line source, E = 1/(4 pi L d) * [s / sqrt(d^2 + s^2)] over the lamp's extent."""
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    sim.init("route")
    sim.set_params(photonCount=1 << 23, maxIterations=2)
    p = sim.params
    d, h, L, lh = 1.0, 1.0, p.lightLength, p.lightHeight
    measured = 123.0
    calibrated = sim.calibrate(measured, h, d)
    # lamp spans [lh, lh + L] above the floor; patch centre at height h, distance d
    s0, s1 = lh - h, lh + L - h
    E = (s1 / np.hypot(d, s1) - s0 / np.hypot(d, s0)) / (4 * np.pi * L * d)
    expect = 0.01 * measured / E
    assert abs(calibrated - expect) / expect < 0.03          # Monte-Carlo error of ~2e4 hits + finite patch
    assert sim.params.lightIntensity == np.float32(calibrated)
    # the room is back: a normal run still matches the oracle
    sim.set_params(photonCount=1 << 20, maxIterations=1, lightIntensity=443.31842)
    seed0 = sim.params.seedState
    dose = sim.run()
    want = oracle_route_run(room, sim.positions, sim.params, 1, sim.params.photonsPerLight, seed0)[2]
    assert dose.tobytes() == want.tobytes()
    sim.close()


def test_shared_reciprocal_division_equals_ieee(ctx):
    samples, bad1, bad2 = ctx.selftest_division(blocks=148 * 16, iters=8192)
    assert samples == 148 * 16 * 256 * 8192
    assert bad2 == 0, "two-step Markstein quotient differs from __fdiv_rn"
    assert bad1 == 0, "one-step Markstein quotient differs from __fdiv_rn"


@pytest.mark.parametrize("host_repack", [0, 1])
def test_error_codes_instead_of_aborts(uv, room, host_repack):
    tris, nodes, tri_idx, _ = room
    c = uv.Context(0)
    c.set_option("host_repack", host_repack)
    with pytest.raises(uv.UvrtError) as e:
        c.extend(10)
    assert e.value.code == -3                                  # no scene yet
    with pytest.raises(uv.UvrtError) as e:
        c.upload_scene(tris, nodes[: 2 * tris.shape[0]], tri_idx)   # the reference's truncated upload (App. B-3)
    assert e.value.code == -1 and "reachable" in str(e.value)
    bad_idx = tri_idx.copy()
    bad_idx[10] = 10**9
    with pytest.raises(uv.UvrtError):
        c.upload_scene(tris, nodes, bad_idx)
    loop = nodes.copy()
    loop[2]["leftFirst"] = 0                                    # a cycle
    with pytest.raises(uv.UvrtError) as e:
        c.upload_scene(tris, loop, tri_idx)
    assert e.value.code == -1 and "twice" in str(e.value)
    span = nodes.copy()
    leaf = int(np.flatnonzero(span["triCount"] > 0)[0])
    span[leaf]["triCount"] = tris.shape[0] + 5                  # a leaf that runs off the end of triIdx
    with pytest.raises(uv.UvrtError) as e:
        c.upload_scene(tris, span, tri_idx)
    assert e.value.code == -1 and "spans" in str(e.value)
    c.upload_scene(tris, nodes, tri_idx)                        # still usable
    c.generate((0, 0, 0), 1.0, 0, 0, 0)                         # zero rays: a no-op
    c.extend(0)
    with pytest.raises(uv.UvrtError):
        c.extend(1 << 40)
    with pytest.raises(uv.UvrtError):
        c.set_option("no_such_option", 1)
    with pytest.raises(uv.UvrtError):
        uv.Context(99)
    c.close()


def _same_tree(a, b):
    (ta, na, ia), (tb, nb, ib) = a, b
    assert np.array_equal(ia, ib), "triIdx differs"
    assert ta.tobytes() == tb.tobytes(), "centroids differ"
    pa, pb = T.reachable_preorder(na), T.reachable_preorder(nb)
    assert np.array_equal(pa, pb), "node numbering differs"
    assert na[pa].tobytes() == nb[pb].tobytes(), "node contents differ"
    assert len(na) == len(nb), "nodesUsed differs"


def test_device_bvh_build_equals_host_builder(uv, ctx, room):
    """uvrt_build_bvh (device) against host/bvh.cpp (which equals the reference's bvh.cpp, test_host.py):
    same triIdx order, same node numbering, same boxes -- on the room, on small and degenerate meshes
    and on a 200k-triangle soup (deeper than the room, several partition patterns)."""
    from importlib import import_module
    B = import_module("small-project-uv-robot-ray-tracer_b200.binding")
    sys_tools = __import__("sys")
    sys_tools.path.insert(0, T.ROOT + "/tools")
    from soup import make_soup
    tris = room[0].copy()
    tris[:, 12:16] = 0
    _same_tree(ctx.build_bvh(tris), B.build_bvh(tris))
    rng = np.random.default_rng(21)
    for n in (1, 2, 3, 5, 17, 64, 200, 257, 5000):
        m = np.zeros((n, 16), dtype=np.float32)
        c = rng.uniform(-5, 5, (n, 3))
        for k in range(3):
            m[:, 4 * k: 4 * k + 3] = (c + rng.uniform(-0.2, 0.2, (n, 3))).astype(np.float32)
        if n >= 17:
            m[5] = m[6]
            m[7, 0:12] = m[7, 0]
            m[9:13] = m[8]             # five identical triangles: a node that cannot be split
        _same_tree(ctx.build_bvh(m), B.build_bvh(m))
    soup = make_soup(200_000)
    _same_tree(ctx.build_bvh(soup), B.build_bvh(soup))
    # a node of 40,000 identical triangles cannot be split: a leaf larger than the multi-block threshold
    same = np.tile(soup[:1], (40_000, 1))
    same = np.concatenate([same, soup[1:3000]])
    _same_tree(ctx.build_bvh(same), B.build_bvh(same))
    # signed zeros at the extrema (common in exported meshes): same tree; a bound may differ in the sign of its zero
    # (include/uvrt.h, uvrt_build_bvh), so compare with the zeros canonicalised
    z = soup[:4000].copy()
    z[::3, 0] = np.float32(-0.0); z[1::3, 0] = np.float32(0.0); z[::5, 5] = np.float32(-0.0); z[2::5, 9] = np.float32(0.0)
    (ta, na, ia), (tb, nb, ib) = ctx.build_bvh(z), B.build_bvh(z)
    assert np.array_equal(ia, ib) and np.array_equal(T.reachable_preorder(na), T.reachable_preorder(nb))
    pa = T.reachable_preorder(na)
    for f in ("min", "max"):
        assert np.array_equal(na[pa][f] + np.float32(0.0), nb[pa][f] + np.float32(0.0))
    assert np.array_equal(na[pa]["leftFirst"], nb[pa]["leftFirst"]) and np.array_equal(na[pa]["triCount"], nb[pa]["triCount"])
    assert np.array_equal(ta + np.float32(0.0), tb + np.float32(0.0))
    # exponentially spaced triangles: every split peels off a few of them, the tree is a deep comb
    deep = soup[:300].copy()
    scale = (np.float32(1.17) ** np.arange(300, dtype=np.float32))[:, None]
    for k in range(3):
        deep[:, 4 * k: 4 * k + 3] *= scale
    td, nd, idd = ctx.build_bvh(deep)
    _same_tree((td, nd, idd), B.build_bvh(deep))
    c2 = uv.Context(0)
    try:
        c2.upload_scene(td, nd, idd)
        assert c2.scene_info()["depth"] < 64
    except uv.UvrtError as e:                     # deeper than the traversal stack: rejected, not mis-traversed
        assert "depth" in str(e)
    c2.close()


def test_raytracer_with_device_built_bvh(uv, room, golden):
    """Mesh::buildBvhOnLoad = false: RayTracer::Init builds the tree with uvrt_build_bvh; the run is
    indistinguishable from one on the host-built tree."""
    sim = uv.Sim(asset_root=T.DATA)
    sim.set_device_bvh(True)
    sim.load_mesh("testroomopt")
    assert sim.mesh_info()["nodesUsed"] == 0
    sim.init("lange_route")
    tris, nodes, tri_idx = sim.mesh_data()
    assert nodes.tobytes() == room[1].tobytes() and np.array_equal(tri_idx, room[2]) and tris.tobytes() == room[0].tobytes()
    sim.set_params(maxIterations=1)
    dose = sim.run()
    assert f"{T.fnv(dose):016x}" == golden["pass_lange_route"]["fnv_dose"]
    sim.close()


def test_checkpoint_resume_is_bit_identical(uv, room, tmp_path):
    """SURVEY section 5 (checkpoint/resume): photon and max maps + counters + the SEED chain make a run
    resumable; two iterations in one go == one iteration, checkpoint, new RayTracer, one more."""
    def new_sim():
        s = uv.Sim(asset_root=T.DATA)
        s.load_mesh("testroomopt")
        s.init("lange_route")
        s.set_params(maxIterations=2, photonCount=1 << 21)
        return s
    a = new_sim()
    whole = a.run()
    a.close()
    b = new_sim()
    b.reset_dosage_map()
    assert b.tick() is False
    ck = tmp_path / "run.ckpt"
    b.save_checkpoint(ck)
    b.close()
    c = new_sim()
    c.reset_dosage_map()
    c.load_checkpoint(ck)
    assert c.tick() is False
    assert c.tick() is True                  # iteration 2 of 2 was the last one
    resumed = c.read_dose()
    assert resumed.tobytes() == whole.tobytes()
    with pytest.raises(uv.UvrtError):
        c.load_checkpoint(tmp_path / "missing.ckpt")
    c.close()


def test_export_dose_ply_json(uv, room, tmp_path):
    """the dose map leaves the process as float32 + PLY with dosageToColor's colours."""
    import json
    s = uv.Sim(asset_root=T.DATA)
    s.load_mesh("testroomopt")
    s.init("route")
    s.set_params(maxIterations=1, photonCount=1 << 20)
    dose = s.run()
    base = tmp_path / "room"
    s.save_dosage_map(base)
    color = s.ctx.read(uv.BUF.COLOR)
    s.close()
    n = room[0].shape[0]
    assert np.fromfile(str(base) + ".dose.f32", dtype=np.float32).tobytes() == dose.tobytes()
    meta = json.load(open(str(base) + ".json"))
    assert meta["triangles"] == n and len(meta["route"]) == 12 and meta["view"] == "dose_mJ_cm2"
    raw = open(str(base) + ".ply", "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex %d" % (3 * n) in head and b"element face %d" % n in head
    vdt = np.dtype([("xyz", "<f4", 3), ("rgb", "u1", 3)])
    v = np.frombuffer(body, dtype=vdt, count=3 * n)
    tris = room[0].reshape(n, 4, 4)[:, :3, :3].reshape(3 * n, 3)
    assert v["xyz"].tobytes() == np.ascontiguousarray(tris).tobytes()
    want = (np.clip(color.reshape(3 * n, 3), 0, 1) * 255 + 0.5).astype(np.uint8)
    assert np.array_equal(v["rgb"], want)
    fdt = np.dtype([("k", "u1"), ("idx", "<i4", 3)])
    f = np.frombuffer(body, dtype=fdt, offset=3 * n * vdt.itemsize, count=n)
    assert (f["k"] == 3).all() and np.array_equal(f["idx"].ravel(), np.arange(3 * n))


def test_device_scene_repack_equals_host_repack(uv, room):
    """uvrt_upload_scene repacks the reference's arrays on the device (csrc/uvrt_scene_prep.cuh); the
    host-side repack stays selectable ("host_repack").  Both must leave the same bytes in HBM: pairs
    in pre-order, triangles in leaf order, same counts, depth and tame flag."""
    from importlib import import_module
    import sys as _sys
    B = import_module("small-project-uv-robot-ray-tracer_b200.binding")
    _sys.path.insert(0, T.ROOT + "/tools")
    from soup import make_soup
    rng = np.random.default_rng(3)
    scenes = [(room[0], room[1], room[2])]
    for n in (1, 2, 3, 17, 257, 5000):
        m = np.zeros((n, 16), dtype=np.float32)
        c = rng.uniform(-5, 5, (n, 3))
        for k in range(3):
            m[:, 4 * k: 4 * k + 3] = (c + rng.uniform(-0.2, 0.2, (n, 3))).astype(np.float32)
        if n >= 17:
            m[9:13] = m[8]                                      # a five-triangle leaf
        scenes.append(B.build_bvh(m))
    scenes.append(B.build_bvh(make_soup(200_000)))
    wild = room[0].copy()
    wild[:, 0:3] *= np.float32(3e6)                              # coordinates outside the tame range
    scenes.append(B.build_bvh(wild))
    images = []
    for mode in (1, 0):
        c = uv.Context(0)
        c.set_option("host_repack", mode)
        out = []
        for tris, nodes, tri_idx in scenes:
            c.upload_scene(tris, nodes, tri_idx)
            info = c.scene_info()
            out.append((info, c.get_option("scene_tame"), c.read(uv.BUF.PAIRS).tobytes(), c.read(uv.BUF.WTRIS).tobytes(),
                        c.scene_upload_bytes()))
        images.append(out)
        c.close()
    for k, (h, d) in enumerate(zip(*images)):
        assert h[0] == d[0], f"scene {k}: scene_info differs: {h[0]} vs {d[0]}"
        assert h[1] == d[1], f"scene {k}: tame flag differs"
        assert h[2] == d[2], f"scene {k}: pairs differ"
        assert h[3] == d[3], f"scene {k}: leaf triangles differ"
        assert d[4] < h[4]                                       # raw arrays are smaller than the repacked image
    assert images[0][-1][1] == 0 and images[0][0][1] == 1


def test_full_default_run_bookkeeping(uv, room):
    """BASELINE configs[1] at full size -- route.xml, 2^25 photons x 10 iterations = 335,544,240 rays, far
    beyond what the oracle traces in seconds -- checked through properties that do not depend on size:
    the maps the RayTracer run leaves behind equal, bit for bit, the f64 sums / maxima formed on the host
    from per-launch integer counts (same SEED chain, launched one by one through the stage API); every
    ray that reports a hit was counted exactly once; the dose inverts back to the photon map.
    (The same run against the golden of the reference's own compiled sources: test_default_run_matches_reference_golden.)"""
    tris, nodes, tri_idx, floor = room
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    sim.init("route")
    p = sim.params
    assert p.photonCount == 1 << 25 and p.maxIterations == 10 and len(sim.positions) == 12
    seed0 = p.seedState
    dose = sim.run()
    c = sim.ctx
    got_sum, got_max = c.read(uv.BUF.SUM), c.read(uv.BUF.MAX)
    assert sim.rays_traced() == 335_544_240 and sim.params.photonMapSize == 335_544_240
    seed_end = sim.params.seedState
    positions = np.array(sim.positions, dtype=np.float32).copy()
    sim.close()

    c2 = uv.Context(0)
    c2.upload_scene(tris, nodes, tri_idx)
    P = int(p.photonsPerLight)
    f32 = np.float32
    want_sum, want_max = np.zeros(tris.shape[0]), np.zeros(tris.shape[0])
    seed, total_hits = seed0, 0
    for it in range(10):
        for k, (x, y, dur) in enumerate(positions):
            lp = (f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y))
            c2.reset(False)
            c2.trace_counts(lp, p.lightLength, 0, P, seed)
            counts = c2.read(uv.BUF.COUNTS)
            if it in (0, 9) and k in (0, 11):
                rays = c2.read(uv.BUF.RAYS, P)        # 89 MB: only a few launches
                hit = rays["dist"] != f32(1e30)
                assert int(hit.sum()) == int(counts.sum())
                assert np.array_equal(np.bincount(rays["triID"][hit], minlength=len(counts)), counts)
            total_hits += int(counts.sum())
            want_sum += counts.astype(np.float64) * np.float64(f32(dur))
            want_max = np.maximum(want_max, counts.astype(np.float64))
            seed = int(c2.seed_chain([lp], p.lightLength, seed)[-1])
    c2.close()
    assert seed == seed_end
    assert got_sum.tobytes() == want_sum.tobytes() and got_max.tobytes() == want_max.tobytes()
    assert 0.85 < total_hits / 335_544_240 < 1.0
    # computeDosage (shade.cl:23-41) inverted: dose * area * photonsPerLight / (I * 0.1) == photon map
    v0, v1, v2 = (tris[:, 0:3].astype(np.float64), tris[:, 4:7].astype(np.float64), tris[:, 8:11].astype(np.float64))
    area = 0.5 * np.linalg.norm(np.cross(v0 - v1, v0 - v2), axis=1)
    ok = area > 1e-12
    back = dose[ok].astype(np.float64) * area[ok] * (335_544_240 // 12) / (np.float64(f32(p.lightIntensity)) * 0.1)
    assert np.allclose(back, want_sum[ok], rtol=2e-5, atol=1e-6)


def test_soup_scene_config5(uv):
    """BASELINE config 5 (incoherent traversal stress), reduced to 1 M triangles: device-built BVH,
    binned and unbinned extend bit-identical to the oracle on the rays the oracle can afford, integer
    counts independent of the ray order, tree deeper than any room (depth > 21)."""
    import sys as _sys
    _sys.path.insert(0, T.ROOT + "/tools")
    from soup import make_soup, soup_route
    c = uv.Context(0)
    n_tris = int(os.environ.get("UVRT_SOUP_TRIS", "1000000"))     # 10000000 = the full config (oracle: ~1 Mrays/s)
    tris, nodes, tri_idx = c.build_bvh(make_soup(n_tris))
    c.upload_scene(tris, nodes, tri_idx)
    info = c.scene_info()
    assert info["inner"] + info["leaves"] <= len(nodes) and info["depth"] > 21 and c.get_option("scene_tame") == 1
    x, z, _ = soup_route()[5]
    lp = (np.float32(x), np.float32(0.5), np.float32(z))
    P = 1_000_000
    counts = []
    for binned in (0, 1):
        c.set_option("bin_rays", binned)
        c.reset(False)
        c.trace_counts(lp, 1.0, 0, P, 3)
        counts.append(c.read(uv.BUF.COUNTS))
    assert np.array_equal(counts[0], counts[1]) and 0.02 < counts[0].sum() / P < 0.9
    n = 100_000
    got = c.read(uv.BUF.RAYS, P)[:n]
    O = T.oracle()
    want = np.zeros(n, dtype=T.RAY_DT)
    O.orc_generate(T.ptr(want), 0, n, lp[0], lp[1], lp[2], np.float32(1.0), 3, None)
    temp = np.zeros(tris.shape[0], dtype=np.int32)
    O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(want), T.ptr(nodes), T.ptr(tri_idx), n, 0, None)
    assert got.tobytes() == want.tobytes()
    c.close()


def test_generate_zero_rng_state_terminates(uv, ctx, room):
    """The reference's generate kernel never returns for a work-item whose seed hashes to 0 (WangHash(61));
    here that work-item keeps its first draw.  Same launch as tests/test_oracle.py's: bit-identical."""
    f32 = np.float32
    lp = (f32(0.0), f32(0.5), f32(np.linspace(-4, 4, 12, dtype=np.float32)[2]))
    n = 4096
    ctx.generate(lp, 1.0, 0, n, 2)
    got = ctx.read(uv.BUF.RAYS, n)
    O = T.oracle()
    want = np.zeros(n, dtype=T.RAY_DT)
    O.orc_generate(T.ptr(want), 0, n, lp[0], lp[1], lp[2], f32(1.0), 2, None)
    assert got.tobytes() == want.tobytes()
    assert got[5]["dir"].tobytes() == np.array([-0.0, -1.0, -0.0], dtype=np.float32).tobytes()
    ctx.reset(False)
    ctx.extend(n)                                   # an axis-parallel ray: literal IEEE path, no hang either
    hit = ctx.read(uv.BUF.RAYS, n)
    tris, nodes, tri_idx, _ = room
    temp = np.zeros(tris.shape[0], dtype=np.int32)
    O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(want), T.ptr(nodes), T.ptr(tri_idx), n, 0, None)
    assert hit.tobytes() == want.tobytes() and np.array_equal(ctx.read(uv.BUF.COUNTS), temp)


def test_headless_driver_cli(uv, golden, tmp_path):
    """uvrt_cli = MyApp::Init + MyApp::Tick without the window (myapp.cpp:15-40, 156-175): one iteration of
    lange_route with the device-built BVH must give the golden dose; --checkpoint/--resume, --export and the
    --json metrics line work from the command line."""
    import json
    import subprocess
    exe = os.path.join(uv.build_dir(), "uvrt_cli")
    base = [exe, "--root", T.DATA, "--room", "testroomopt", "--route", "lange_route"]
    out = tmp_path / "dose.f32"
    r = subprocess.run(base + ["--iterations", "1", "--device-bvh", "--out", str(out), "--export", str(tmp_path / "room"), "--json"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    dose = np.fromfile(out, dtype=np.float32)
    assert f"{T.fnv(dose):016x}" == golden["pass_lange_route"]["fnv_dose"]
    assert "Progress: 100%" in r.stdout and "device build" in r.stdout
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["rays"] == 12 * 2796202 and line["triangles"] == 44866 and line["stage_ms"]["extend"] > 0
    assert (tmp_path / "room.ply").exists() and (tmp_path / "room.json").exists()
    assert np.fromfile(tmp_path / "room.dose.f32", dtype=np.float32).tobytes() == dose.tobytes()
    # two iterations at once == one iteration + checkpoint, then a resumed process for the second
    two = tmp_path / "two.f32"
    ck = tmp_path / "run.ckpt"
    resumed = tmp_path / "resumed.f32"
    for args in (["--iterations", "2", "--out", str(two)],
                 ["--iterations", "1", "--checkpoint", str(ck)],
                 ["--iterations", "2", "--resume", str(ck), "--out", str(resumed)]):
        r = subprocess.run(base + ["--photons", str(1 << 21)] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert np.fromfile(two, dtype=np.float32).tobytes() == np.fromfile(resumed, dtype=np.float32).tobytes()
    bad = subprocess.run(base[:3] + ["--room", "no_such_room"], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 1 and "cannot load room" in bad.stderr


def test_stage_calls_and_trace_share_the_count_buffer(uv, ctx, room):
    """generate + extend through the stage API leave their counts for the next accumulate, and uvrt_trace's
    accumulate must see them even though uvrt_trace normally alternates between two count buffers
    (reference order: extend, extend, accumulate adds both launches' photons with the second duration)."""
    lp = lange_pos0(room[3])
    P = 200_000
    for start_slot in (0, 1):
        ctx.reset(True)
        if start_slot:
            ctx.trace(lp, 1.0, 1.0, 0, 70_000, 5)            # moves the context to the other ray slot
            ctx.reset(True)
        ctx.trace_counts(lp, 1.0, 0, P, 11)                   # counts only
        first = ctx.read(uv.BUF.COUNTS).astype(np.float64)
        ctx.trace(lp, 1.0, 2.5, 0, P, 12)                     # extend + accumulate
        total = ctx.read(uv.BUF.SUM)
        assert not ctx.read(uv.BUF.COUNTS).any()              # accumulate zeroed what it consumed
        ctx.reset(True)
        ctx.trace_counts(lp, 1.0, 0, P, 12)
        second = ctx.read(uv.BUF.COUNTS).astype(np.float64)
        assert np.array_equal(total, (first + second) * 2.5)
    ctx.reset(True)


# ---- reference-derived goldens of the benchmarked configurations (tests/golden/make_golden_runs.py) ----------
def _run_golden(route):
    import json
    return json.load(open(os.path.join(T.ROOT, "tests", "golden", "route_runs.json")))["runs"][route]


@pytest.mark.parametrize("route", ["route", "lange_route"])
def test_default_run_matches_reference_golden(uv, room, route):
    """BASELINE configs[1] (route.xml) and the config-2 substitute (lange_route.xml) at FULL size: 2^25 photons x
    10 iterations = 120 launches of 2,796,202 rays = 335,544,240 rays through RayTracer, compared after every
    iteration with what the reference's own compiled sources produced (photon map, max map, dose, colours, SEED)."""
    g = _run_golden(route)
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    sim.init(route)
    p = sim.params
    assert p.photonCount == 1 << 25 and p.maxIterations == 10 and p.photonsPerLight == g["photonsPerLight"]
    assert int(np.float32(p.lightIntensity).view(np.uint32)) == g["lightIntensity_bits"]
    sim.reset_dosage_map()
    c = sim.ctx
    for it in range(10):
        finished = sim.tick()
        a = g["after_iteration"][it]
        got = (f"{T.fnv(c.read(uv.BUF.SUM)):016x}", f"{T.fnv(c.read(uv.BUF.MAX)):016x}",
               f"{T.fnv(sim.read_dose()):016x}", f"{T.fnv(c.read(uv.BUF.COLOR)):016x}", sim.params.seedState)
        assert got == (a["fnv_photonMap"], a["fnv_maxPhotonMap"], a["fnv_dose"], a["fnv_color"], a["seed"]), f"iteration {it + 1}"
    assert finished is False and sim.tick() is True          # the eleventh tick only notices that the run is over
    assert sim.rays_traced() == g["rays"] == 335_544_240
    dose = sim.read_dose()
    assert [int(v) for v in dose[:8].view(np.uint32)] == g["after_iteration"][9]["dose_head_bits"]
    sim.close()


def test_every_launch_of_the_default_run_matches_reference_counts(uv, ctx, room):
    """The 120 count vectors of the default run (route.xml), launch by launch through the stage API with the
    golden SEED chain: hits and FNV of the integer counts equal the compiled reference's for every launch."""
    g = _run_golden("route")
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_route("route")
    pos, p = sim.positions, sim.params
    sim.close()
    f32 = np.float32
    P = g["photonsPerLight"]
    for k in range(120):
        x, y, _ = pos[k % 12]
        lp = (f32(x), f32(f32(room[3]) + f32(p.lightHeight)), f32(y))
        ctx.reset(False)
        ctx.trace_counts(lp, p.lightLength, 0, P, g["seed_chain"][k])
        counts = ctx.read(uv.BUF.COUNTS)
        assert int(counts.sum()) == g["hits_per_launch"][k], f"launch {k}"
        assert f"{T.fnv(counts):016x}" == g["fnv_counts_per_launch"][k], f"launch {k}"
    ctx.reset(True)


@pytest.mark.parametrize("n_tris", [1_000_000, 10_000_000])
def test_soup_matches_reference_golden(uv, n_tris):
    """BASELINE config 5 at full size (10 M triangles) and at 1 M: device-built BVH, 1e6-ray launches; triIdx,
    per-ray (dist, triID) and integer counts hash-equal to the reference's builder + extend kernel
    (tests/golden/soup.json)."""
    import json
    import sys as _sys
    _sys.path.insert(0, T.ROOT + "/tools")
    from soup import make_soup, soup_route
    g = json.load(open(os.path.join(T.ROOT, "tests", "golden", "soup.json")))["scenes"][str(n_tris)]
    c = uv.Context(0)
    tris, nodes, tri_idx = c.build_bvh(make_soup(n_tris))
    assert f"{T.fnv(tri_idx):016x}" == g["fnv_triIdx"]
    c.upload_scene(tris, nodes, tri_idx)
    del tris, nodes
    route = soup_route()
    for key, gl in g["launch"].items():
        x, z, _ = route[gl["position"]]
        lp = (np.float32(x), np.float32(0.5), np.float32(z))
        c.reset(False)
        c.trace_counts(lp, 1.0, 0, gl["rays"], gl["seed_in"])
        rays = c.read(uv.BUF.RAYS, gl["rays"])
        counts = c.read(uv.BUF.COUNTS)
        assert int(counts.sum()) == gl["hits"], key
        assert f"{int(T.oracle().orc_fnv_hits(T.ptr(rays), gl['rays'])):016x}" == gl["fnv_hits"], key
        assert f"{T.fnv(counts):016x}" == gl["fnv_counts"], key
        assert int(c.seed_chain([lp], 1.0, gl["seed_in"])[1]) == gl["seed_out"]
    c.close()


def test_cuda_path_against_the_compiled_reference_directly(uv, ctx, room):
    """Where oracle/_ref/libuvrt_ref.so travelled to the box: one launch compared with the reference's own
    kernel sources directly (not through the C port): rays, closest hits, counts, maps, dose, colours."""
    if not T.ref_available():
        pytest.skip("oracle/_ref/libuvrt_ref.so is not on this box")
    R = T.ref()
    tris, nodes, tri_idx, floor = room
    n, P, f32 = tris.shape[0], 400_000, np.float32
    lp = (f32(0.51000142), f32(f32(floor) + f32(0.6)), f32(-0.25500044))
    want = np.zeros(P, dtype=T.RAY_DT)
    so = C.c_uint(0)
    R.ref_generate(T.ptr(want), 0, P, lp[0], lp[1], lp[2], f32(1.0), 99, C.byref(so))
    ctx.reset(True)
    ctx.generate(lp, 1.0, 0, P, 99)
    assert ctx.read(uv.BUF.RAYS, P).tobytes() == want.tobytes()
    temp = np.zeros(n, dtype=np.int32)
    R.ref_extend(T.ptr(temp), T.ptr(tris), T.ptr(want), T.ptr(nodes), T.ptr(tri_idx), P, n, 0)
    ctx.extend(P)
    assert ctx.read(uv.BUF.RAYS, P).tobytes() == want.tobytes()
    assert np.array_equal(ctx.read(uv.BUF.COUNTS), temp)
    pm, mx = np.zeros(n), np.zeros(n)
    R.ref_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), f32(37.5), n)
    ctx.accumulate(37.5)
    assert ctx.read(uv.BUF.SUM).tobytes() == pm.tobytes() and ctx.read(uv.BUF.MAX).tobytes() == mx.tobytes()
    dose, col = np.zeros(n, dtype=np.float32), np.zeros((n, 9), dtype=np.float32)
    R.ref_compute_dosage(T.ptr(pm), T.ptr(dose), T.ptr(tris), P, f32(44.331842), n)
    R.ref_dosage_to_color(T.ptr(dose), T.ptr(col), f32(100.0), 0, n)
    ctx.shade(0, P, 44.331842)
    ctx.color(100.0, 0)
    assert ctx.read(uv.BUF.DOSE).tobytes() == dose.tobytes() and ctx.read(uv.BUF.COLOR).tobytes() == col.tobytes()
    assert int(ctx.seed_chain([lp], 1.0, 99)[1]) == so.value
    ctx.reset(True)


def test_count_matrix_ray_ranges_fold_to_single_context_result(uv, ctx, room):
    """launches cut into ray ranges over two 'ranks' (two contexts on one GPU), integer count
    rows exchanged and summed by the caller (UVRT_BUF_MATRIX; on real ranks uvrt_matrix_fold does it with one
    ncclAllReduce), folded in launch order: photon map AND per-launch max map bit-identical to one context running
    uvrt_trace launch by launch -- with durations whose products do not add exactly."""
    f32 = np.float32
    tris, nodes, tri_idx, floor = room
    lamps = [(f32(-0.255), f32(floor + f32(0.6)), f32(-3.31)), (f32(0.085), f32(floor + f32(0.6)), f32(-2.46)),
             (f32(-0.51), f32(floor + f32(0.6)), f32(-1.19)), (f32(0.3), f32(floor + f32(0.6)), f32(2.0))]
    durs = np.array([0.1, 7.3, 60.0, 1e-3], dtype=np.float32)
    P = 300_001
    seeds = [0]
    for lp in lamps:
        seeds.append(int(ctx.seed_chain([lp], 1.0, seeds[-1])[1]))
    ctx.reset(True)
    for lp, dur, seed in zip(lamps, durs, seeds):
        ctx.trace(lp, 1.0, dur, 0, P, seed)
    want_sum, want_max = ctx.read(uv.BUF.SUM), ctx.read(uv.BUF.MAX)
    cuts = [0, 70_001, 70_001, 200_000, P]                     # ranges of unequal size, one of them empty
    ranks = [uv.Context(0), uv.Context(0)]
    mats = []
    for r, c in enumerate(ranks):
        c.upload_scene(tris, nodes, tri_idx)
        c.reset(True)
        c.matrix_begin(len(lamps))
        for row, (lp, seed) in enumerate(zip(lamps, seeds)):
            for j in range(len(cuts) - 1):
                if (row + j) % 2 == r and cuts[j + 1] > cuts[j]:
                    c.trace_row(row, lp, 1.0, cuts[j], cuts[j + 1] - cuts[j], seed)
        mats.append(c.read(uv.BUF.MATRIX, len(lamps)))
    total = mats[0] + mats[1]
    for c in ranks:
        c.write(uv.BUF.MATRIX, total)
        c.matrix_fold(durs, reduce=False)
        assert c.read(uv.BUF.SUM).tobytes() == want_sum.tobytes()
        assert c.read(uv.BUF.MAX).tobytes() == want_max.tobytes()
        c.close()
    assert mats[0].any() and mats[1].any()
    ctx.reset(True)


def test_host_seed_chain_equals_device_seed_chain(uv, ctx, room):
    """RayTracer advances SEED on the host (RayTracer::SeedAfterLaunch); uvrt_seed_chain replays work-item 0 on the device."""
    H = uv.host()
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_route("lange_route")
    pos, p = sim.positions, sim.params
    sim.close()
    f32 = np.float32
    lps = [(f32(x), f32(f32(room[3]) + f32(p.lightHeight)), f32(y)) for x, y, _ in pos] * 3 + [(f32(0.0), f32(0.5), f32(-2.5454545))]
    for seed0 in (0, 2, 0xfffffff0):
        chain = ctx.seed_chain(lps, p.lightLength, seed0)
        seed = seed0
        for k, lp in enumerate(lps):
            seed = int(H.uvrt_host_seed_after_launch(lp[0], lp[1], lp[2], f32(p.lightLength), seed))
            assert seed == int(chain[k + 1])


def test_fast_extend_default_run_matches_reference_golden(uv, room, variant=50):
    """The certified fast extend on BASELINE configs[1] at full size (335,544,240 rays): every map of the run hashes
    to the golden of the reference's own compiled sources, i.e. not one ray of the run got another triangle or
    distance; and with "fast_check" (every ray also traced in reference order) no certified ray disagrees."""
    g = _run_golden("route")
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    sim.init("route")
    c = sim.ctx
    assert c.get_option("scene_nested") == 1 and c.get_option("fast_ready") == 1
    c.set_option("extend_variant", variant)
    c.fast_stats(reset=True)
    dose = sim.run()
    a = g["after_iteration"][9]
    got = (f"{T.fnv(c.read(uv.BUF.SUM)):016x}", f"{T.fnv(c.read(uv.BUF.MAX)):016x}", f"{T.fnv(dose):016x}", f"{T.fnv(c.read(uv.BUF.COLOR)):016x}")
    assert got == (a["fnv_photonMap"], a["fnv_maxPhotonMap"], a["fnv_dose"], a["fnv_color"])
    st = c.fast_stats(reset=True)
    assert st["ineligible"] < 335_544_240 * 1e-4          # axis-parallel rays and the like
    assert st["cert_fallbacks"] < 335_544_240 * 1e-3      # two surfaces within 2^-12: re-traced in reference order
    # one pass with every ray traced both ways
    c.set_option("fast_check", 1)
    sim.set_params(maxIterations=1)
    sim.set_seed(0)
    sim.run()
    st = c.fast_stats(reset=True)
    assert st["check_mismatches"] == 0, st
    assert f"{T.fnv(sim.read_dose()):016x}" == g["after_iteration"][0]["fnv_dose"]
    sim.close()


def test_fast_extend_soup_check_mode(uv):
    """The 1 M-triangle soup (deep, incoherent traversal) through the fast kernel with every ray traced both
    ways: no certified ray disagrees with the reference order, and the counts equal the golden."""
    import json
    import sys as _sys
    _sys.path.insert(0, T.ROOT + "/tools")
    from soup import make_soup, soup_route
    g = json.load(open(os.path.join(T.ROOT, "tests", "golden", "soup.json")))["scenes"]["1000000"]["launch"]["5:3"]
    c = uv.Context(0)
    tris, nodes, tri_idx = c.build_bvh(make_soup(1_000_000))
    c.upload_scene(tris, nodes, tri_idx)
    assert c.get_option("fast_ready") == 1
    x, z, _ = soup_route()[5]
    lp = (np.float32(x), np.float32(0.5), np.float32(z))
    assert c.get_option("extend_variant") == 50              # the default wherever the fast kernel can serve the scene
    for variant in (50,):
        c.set_option("extend_variant", variant)
        c.set_option("fast_check", 1)
        c.fast_stats(reset=True)
        c.reset(False)
        c.trace_counts(lp, 1.0, 0, g["rays"], g["seed_in"])
        counts = c.read(uv.BUF.COUNTS)
        rays = c.read(uv.BUF.RAYS, g["rays"])
        st = c.fast_stats(reset=True)
        assert st["check_mismatches"] == 0, (variant, st)
        assert f"{T.fnv(counts):016x}" == g["fnv_counts"]
        assert f"{int(T.oracle().orc_fnv_hits(T.ptr(rays), g['rays'])):016x}" == g["fnv_hits"]
        c.set_option("fast_check", 0)
        c.reset(False)
        c.trace_counts(lp, 1.0, 0, g["rays"], g["seed_in"])
        assert f"{T.fnv(c.read(uv.BUF.COUNTS)):016x}" == g["fnv_counts"]
    c.close()


def test_fast_extend_falls_back_on_scenes_it_cannot_serve(uv, room):
    """Boxes that are not nested (a hand-made tree) or coordinates outside the tame range: variant 50 must take the
    exact kernel and still give the exact kernel's answer."""
    tris, nodes, tri_idx, floor = room
    bad = nodes.copy()
    k = int(np.flatnonzero(bad["triCount"] == 0)[5])
    bad[k]["min"] = bad[k]["min"] + np.float32(0.01)          # the node's box no longer contains its children
    c = uv.Context(0)
    c.upload_scene(tris, bad, tri_idx)
    assert c.get_option("scene_nested") == 0 and c.get_option("fast_ready") == 0
    lp = lange_pos0(floor)
    out = []
    for variant in (2, 50):
        c.set_option("extend_variant", variant)
        c.reset(True)
        c.trace_counts(lp, 1.0, 0, 200_000, 5)
        out.append((c.read(uv.BUF.RAYS, 200_000).tobytes(), c.read(uv.BUF.COUNTS).tobytes()))
    assert out[0] == out[1]
    c.close()


def test_cost_probe_is_deterministic_and_tracks_the_oracle_counters(uv, ctx, room):
    """uvrt_probe_cost (the measurement behind the cost-aware deal): the same numbers on every context, equal to the
    oracle's traversal counters on the same rays, and different enough between lamp positions to matter."""
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_route("route")
    pos, p = sim.positions, sim.params
    sim.close()
    f32 = np.float32
    tris, nodes, tri_idx, floor = room
    other = uv.Context(0)
    other.upload_scene(tris, nodes, tri_idx)
    costs = []
    for k in (0, 5, 11):
        lp = (f32(pos[k, 0]), f32(f32(floor) + f32(p.lightHeight)), f32(pos[k, 1]))
        a = ctx.probe_cost(lp, p.lightLength, 0, 8192)
        assert a == other.probe_cost(lp, p.lightLength, 0, 8192)
        rays = np.zeros(8192, dtype=T.RAY_DT)
        T.oracle().orc_generate(T.ptr(rays), 0, 8192, lp[0], lp[1], lp[2], f32(p.lightLength), 0, None)
        cnt = T.Counters()
        temp = np.zeros(tris.shape[0], dtype=np.int32)
        T.oracle().orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), 8192, 0, C.byref(cnt))
        assert a == (cnt.innerVisits / 8192, cnt.triTests / 8192)
        costs.append(44 * a[0] + 75 * a[1])
    other.close()
    assert max(costs) / min(costs) > 1.1


def test_timeline_log_of_a_run(uv, room, tmp_path):
    """The "timeline" option (tracing): every C-ABI call with its host time and every stage launch with
    its enqueue time and device start/stop, as JSON; NVTX ranges ride on the same scopes."""
    import json
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    sim.init("route")
    sim.set_params(photonCount=1 << 21, maxIterations=1)
    c = sim.ctx
    c.set_option("timeline", 1)
    sim.run()
    path = tmp_path / "timeline.json"
    c.timeline_dump(path)
    c.set_option("timeline", 0)
    t = json.load(open(path))
    names = [k[0] for k in t["kernels"]]
    assert names.count("extend") == 12 and names.count("generate") == 12 and "computeDosage" in names and "reset" in names
    for name, host_us, start_us, stop_us in t["kernels"]:
        assert stop_us >= start_us >= 0 and host_us >= 0
    calls = [k[0] for k in t["calls"]]
    assert calls.count("uvrt_trace") == 12 and "uvrt_reset" in calls and "uvrt_read" in calls
    sim.close()
