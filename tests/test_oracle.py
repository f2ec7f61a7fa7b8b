"""The oracle (oracle/uvrt_oracle.c, a plain-C port) pinned against
  * the golden vectors generated from the reference's own compiled sources
    (tests/golden/appendix_c.json), and
  * the reference's own compiled sources directly (oracle/_ref), when they were built here.
CPU only."""
import ctypes as C

import numpy as np
import pytest

import uvrt_testlib as T
from conftest import lange_pos0


def _launch(lib_gen, lib_ext, is_ref, tris, nodes, tri_idx, lp, P, seed_in=0, first=0):
    rays = np.zeros(P, dtype=T.RAY_DT)
    so = (C.c_uint if is_ref else C.c_uint32)(0)
    lib_gen(T.ptr(rays), first, P, lp[0], lp[1], lp[2], 1.0, seed_in, C.byref(so))
    gen = rays.copy()
    temp = np.zeros(tris.shape[0], dtype=np.int32)
    if is_ref:
        lib_ext(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, tris.shape[0], 0)
    else:
        lib_ext(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, 0, None)
    return gen, rays, temp, int(so.value)


def test_wang_hash_and_float_conversion(checkers):
    O = checkers.oracle()
    # tools.cl:2 written out by hand for s = 1
    s = 1
    s = (s ^ 61) ^ (s >> 16); s = (s * 9) & 0xffffffff; s ^= s >> 4; s = (s * 0x27d4eb2d) & 0xffffffff; s ^= s >> 15
    assert O.orc_wang_hash(1) == s


@pytest.mark.parametrize("P", [1000000, 2796202])
def test_oracle_matches_golden_launch(checkers, room, golden, P):
    tris, nodes, tri_idx, floor = room
    g = golden["launch"][str(P)]
    lp = lange_pos0(floor)
    assert [int(np.float32(v).view(np.uint32)) for v in lp] == golden["lightPos_bits"]
    O = checkers.oracle()
    gen, rays, temp, seed_out = _launch(O.orc_generate, O.orc_extend, False, tris, nodes, tri_idx, lp, P)
    assert seed_out == g["seed_out"]
    assert f"{T.fnv(gen):016x}" == g["fnv_rays_after_generate"]
    assert f"{int(O.orc_fnv_hits(T.ptr(rays), P)):016x}" == g["fnv_hits"]
    assert f"{T.fnv(temp):016x}" == g["fnv_counts"]
    assert int(temp.sum()) == g["hits"]
    assert [int(temp.argmax()), int(temp.max())] == g["hottest"]
    for i, hexs in g["samples"].items():
        assert rays[int(i)].tobytes().hex() == hexs
    # threads 0..2 share a ray: their seed expression is negative and saturates (App. B-2)
    assert gen[0].tobytes() == gen[1].tobytes() == gen[2].tobytes() != gen[3].tobytes()


def test_oracle_matches_compiled_reference(checkers, room):
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    tris, nodes, tri_idx, floor = room
    O, R = checkers.oracle(), checkers.ref()
    rng = np.random.default_rng(7)
    for trial in range(3):
        lp = (np.float32(rng.uniform(-1.2, 0.5)), np.float32(floor + rng.uniform(0.2, 0.9)), np.float32(rng.uniform(-3.3, 4.0)))
        seed_in = int(rng.integers(0, 2**32))
        first = int(rng.integers(0, 5_000_000)) if trial else 0
        P = 200_000
        a = _launch(O.orc_generate, O.orc_extend, False, tris, nodes, tri_idx, lp, P, seed_in, first)
        b = _launch(R.ref_generate, R.ref_extend, True, tris, nodes, tri_idx, lp, P, seed_in, first)
        assert a[0].tobytes() == b[0].tobytes(), "generate differs"
        assert a[1].tobytes() == b[1].tobytes(), "hits differ"
        assert np.array_equal(a[2], b[2])
        assert a[3] == b[3]


def test_per_triangle_passes_match_compiled_reference(checkers, room):
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not built")
    tris, _, _, _ = room
    n = tris.shape[0]
    O, R = checkers.oracle(), checkers.ref()
    rng = np.random.default_rng(3)
    temp = rng.integers(0, 60000, n).astype(np.int32)
    temp[::7] = 0
    outs = []
    for lib, pre in ((O, "orc"), (R, "ref")):
        pm = rng.random(n) * 0 + np.arange(n) * 0.25
        mx = np.full(n, 100.0)
        t = temp.copy()
        getattr(lib, pre + "_accumulate")(T.ptr(pm), T.ptr(mx), T.ptr(t), np.float32(0.1), n)
        getattr(lib, pre + "_accumulate")(T.ptr(pm), T.ptr(mx), T.ptr(temp.copy()), np.float32(60), n)
        dose = np.zeros(n, dtype=np.float32)
        getattr(lib, pre + "_compute_dosage")(T.ptr(pm), T.ptr(dose), T.ptr(tris), 2796202, np.float32(44.019705), n)
        col = np.zeros((n, 9), dtype=np.float32)
        col2 = np.zeros((n, 9), dtype=np.float32)
        getattr(lib, pre + "_dosage_to_color")(T.ptr(dose), T.ptr(col), np.float32(100), 0, n)
        getattr(lib, pre + "_dosage_to_color")(T.ptr(dose), T.ptr(col2), np.float32(300), 1, n)
        outs.append((pm, mx, t, dose, col, col2))
    for a, b in zip(*outs):
        assert a.tobytes() == b.tobytes()
    assert not outs[0][2].any()          # temp map zeroed


@pytest.mark.slow
def test_oracle_full_pass_matches_golden(checkers, room, golden, uv):
    """One pass over lange_route (12 x 2,796,202 rays): SEED chain, hits, dose (App. C.2)."""
    tris, nodes, tri_idx, floor = room
    g = golden["pass_lange_route"]
    O = checkers.oracle()
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_route("lange_route")
    pos, p = sim.positions, sim.params
    P, n = int(p.photonsPerLight), tris.shape[0]
    assert P == g["photonsPerLight"]
    f32 = np.float32
    pm, mx, temp = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    seed, chain, hits = 0, [0], []
    rays = np.zeros(P, dtype=T.RAY_DT)
    for (x, y, dur) in pos:
        so = C.c_uint32(0)
        O.orc_generate(T.ptr(rays), 0, P, f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y), f32(p.lightLength), seed, C.byref(so))
        O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, 0, None)
        hits.append(int(temp.sum()))
        O.orc_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), f32(dur), n)
        seed = int(so.value)
        chain.append(seed)
    assert chain == g["seed_chain"]
    assert hits == g["hits_per_position"]
    dose = np.zeros(n, dtype=np.float32)
    O.orc_compute_dosage(T.ptr(pm), T.ptr(dose), T.ptr(tris), P, f32(f32(p.lightIntensity) * f32(0.1)), n)
    assert f"{T.fnv(pm):016x}" == g["fnv_photonMap"]
    assert f"{T.fnv(mx):016x}" == g["fnv_maxPhotonMap"]
    assert f"{T.fnv(dose):016x}" == g["fnv_dose"]
    col = np.zeros((n, 9), dtype=np.float32)
    O.orc_dosage_to_color(T.ptr(dose), T.ptr(col), f32(p.minDosage), 0, n)
    assert f"{T.fnv(col):016x}" == g["fnv_color"]
    assert int((dose == 0).sum()) == g["unlit"]


def test_oracle_first_route_pass_matches_run_golden(checkers, room, uv):
    """BASELINE configs[1] itself: the first of the ten passes over the shipped route.xml (12 x 2,796,202 rays,
    lamp height 0.6, power 443.3) against tests/golden/route_runs.json, which the reference's own compiled
    sources produced for the whole 10-iteration run (tests/golden/make_golden_runs.py)."""
    import json
    import os
    g = json.load(open(os.path.join(T.ROOT, "tests", "golden", "route_runs.json")))["runs"]["route"]
    tris, nodes, tri_idx, floor = room
    O = checkers.oracle()
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_route("route")
    pos, p = sim.positions, sim.params
    sim.close()
    P, n, f32 = int(p.photonsPerLight), tris.shape[0], np.float32
    assert P == g["photonsPerLight"] and len(pos) == g["positions"]
    assert int(f32(p.lightHeight).view(np.uint32)) == g["lightHeight_bits"]
    pm, mx, temp = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
    rays = np.zeros(P, dtype=T.RAY_DT)
    seed = 0
    for k, (x, y, dur) in enumerate(pos):
        so = C.c_uint32(0)
        O.orc_generate(T.ptr(rays), 0, P, f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y), f32(p.lightLength), seed, C.byref(so))
        O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, 0, None)
        assert int(temp.sum()) == g["hits_per_launch"][k]
        assert f"{T.fnv(temp):016x}" == g["fnv_counts_per_launch"][k]
        O.orc_accumulate(T.ptr(pm), T.ptr(mx), T.ptr(temp), f32(dur), n)
        seed = int(so.value)
        assert seed == g["seed_chain"][k + 1]
    a = g["after_iteration"][0]
    dose = np.zeros(n, dtype=np.float32)
    O.orc_compute_dosage(T.ptr(pm), T.ptr(dose), T.ptr(tris), P, f32(f32(p.lightIntensity) * f32(0.1)), n)
    col = np.zeros((n, 9), dtype=np.float32)
    O.orc_dosage_to_color(T.ptr(dose), T.ptr(col), f32(p.minDosage), 0, n)
    assert (f"{T.fnv(pm):016x}", f"{T.fnv(mx):016x}", f"{T.fnv(dose):016x}", f"{T.fnv(col):016x}") == \
           (a["fnv_photonMap"], a["fnv_maxPhotonMap"], a["fnv_dose"], a["fnv_color"])


def test_run_goldens_agree_with_appendix_c(golden):
    """The 10-iteration golden of lange_route starts with the pass SURVEY App. C.2 records."""
    import json
    import os
    g = json.load(open(os.path.join(T.ROOT, "tests", "golden", "route_runs.json")))["runs"]["lange_route"]
    c2 = golden["pass_lange_route"]
    assert g["seed_chain"][:13] == c2["seed_chain"] and g["hits_per_launch"][:12] == c2["hits_per_position"]
    a = g["after_iteration"][0]
    assert (a["fnv_photonMap"], a["fnv_maxPhotonMap"], a["fnv_dose"], a["fnv_color"]) == \
           (c2["fnv_photonMap"], c2["fnv_maxPhotonMap"], c2["fnv_dose"], c2["fnv_color"])


def test_bvh_closest_hit_equals_brute_force(checkers, room):
    """Slab-test false negatives would show up as BVH hits farther than the brute-force hit."""
    tris, nodes, tri_idx, floor = room
    O = checkers.oracle()
    P = 4000
    rays = np.zeros(P, dtype=T.RAY_DT)
    lp = lange_pos0(floor)
    O.orc_generate(T.ptr(rays), 0, P, lp[0], lp[1], lp[2], 1.0, 123, None)
    brute = rays.copy()
    temp = np.zeros(tris.shape[0], dtype=np.int32)
    O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(rays), T.ptr(nodes), T.ptr(tri_idx), P, 0, None)
    O.orc_brute_force(T.ptr(tris), tris.shape[0], T.ptr(brute), P)
    assert np.array_equal(rays["dist"], brute["dist"])
    same = rays["triID"] == brute["triID"]
    # equal distance but a different ID is an exact tie broken by visiting order
    assert np.all(same | (rays["dist"] == brute["dist"]))
    assert same.mean() > 0.999


def test_reset(checkers):
    O = checkers.oracle()
    n = 100
    pm, mx, t, col = np.ones(n), np.ones(n), np.ones(n, dtype=np.int32), np.ones((n, 9), dtype=np.float32)
    O.orc_reset(T.ptr(pm), T.ptr(mx), T.ptr(t), T.ptr(col), 0, n)
    assert not pm.any() and not mx.any() and not t.any() and col.all()
    O.orc_reset(T.ptr(pm), T.ptr(mx), T.ptr(t), T.ptr(col), 1, n)
    assert not col.any()


def test_zero_rng_state_does_not_hang(checkers):
    """WangHash(61) == 0 and xorshift32 maps 0 to 0: a work-item whose seed expression evaluates to 61 draws
    zeros for ever and the reference's rejection loop (generate.cl:25-28) never terminates -- a GPU hang in
    the reference (found by the 10 M-triangle sweep: lamp (0, 0.5, -2.5454545), SEED 2, work-item 5).  The
    port and the CUDA kernel keep the first draw instead: origin at the lamp's foot, direction straight down."""
    O = T.oracle()
    assert O.orc_wang_hash(61) == 0
    f32 = np.float32
    lp = (f32(0.0), f32(0.5), f32(np.linspace(-4, 4, 12, dtype=np.float32)[2]))
    rays = np.zeros(16, dtype=T.RAY_DT)
    so = C.c_uint32(7)
    O.orc_generate(T.ptr(rays), 0, 16, lp[0], lp[1], lp[2], f32(1.0), 2, C.byref(so))
    r = rays[5]
    assert r["dir"].tobytes() == np.array([-0.0, -1.0, -0.0], dtype=np.float32).tobytes()
    assert r["orig"].tobytes() == np.array(lp, dtype=np.float32).tobytes()
    assert r["dist"] == f32(1e30) and r["triID"] == 0
    assert np.all(np.abs(np.linalg.norm(rays["dir"].astype(np.float64), axis=1) - 1.0) < 1e-6)
    # work-item 0 stuck at zero: SEED_out is 0 (lamp term 60: 0*17 + 1 + 60 = 61)
    so = C.c_uint32(7)
    O.orc_generate(T.ptr(rays), 0, 1, f32(0.0), f32(0.0), f32(60.0 / 11.0), f32(1.0), 0, C.byref(so))
    e = f32(1.0) + f32(0.0) * f32(13.0)
    e = f32(e + f32(0.0) * f32(7.0))
    e = f32(e + f32(f32(60.0 / 11.0) * f32(11.0)))
    if int(e) == 61:
        assert so.value == 0
