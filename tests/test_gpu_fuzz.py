"""Randomised cross-checks on awkward meshes (planar, collinear, clustered, duplicated, zero-area, tiny and huge
extents, grids of equal centroids): the device BVH build must equal the host builder, the device scene repack the
host repack, and extend the oracle -- all byte for byte.  Coordinates outside the tame range send every ray down
the literal IEEE-division path, so this is also where that path meets real trees."""
import ctypes as C
import importlib

import numpy as np
import pytest

import uvrt_testlib as T

pytestmark = pytest.mark.gpu

KINDS = ["uniform", "planar", "line", "clusters", "tiny", "huge", "grid", "dups", "degenerate"]


def make_mesh(rng, n, kind):
    c = rng.uniform(-5, 5, (n, 3))
    ext = 0.2
    if kind == "planar":
        c[:, rng.integers(3)] = rng.uniform(-1, 1)
    elif kind == "line":
        c[:, 1:] = c[:1, 1:]
    elif kind == "clusters":
        k = rng.integers(1, 6)
        c = rng.normal(0, 0.01, (n, 3)) + rng.uniform(-5, 5, (k, 3))[rng.integers(k, size=n)]
        ext = 0.001
    elif kind == "tiny":
        c *= 1e-18
        ext = 1e-20
    elif kind == "huge":
        c *= 1e14
        ext = 1e12
    elif kind == "grid":
        c = np.round(c)                      # many equal centroids and equal bin coordinates
        ext = 0.0
    m = np.zeros((n, 16), dtype=np.float32)
    for k in range(3):
        m[:, 4 * k: 4 * k + 3] = (c + rng.uniform(-ext, ext, (n, 3))).astype(np.float32)
    if kind == "grid":
        m[:, 4] += np.float32(0.5)
        m[:, 9] += np.float32(0.5)
    if kind == "dups" and n > 4:
        m[rng.integers(n, size=n // 2)] = m[rng.integers(n, size=n // 2)]
    if kind == "degenerate":
        m[:, 4:7] = m[:, 0:3]                # zero-area triangles
    return m


@pytest.fixture(scope="module")
def uv():
    return importlib.import_module("small-project-uv-robot-ray-tracer_b200")


def test_build_repack_extend_on_awkward_meshes(uv):
    B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
    rng = np.random.default_rng(1)
    O = T.oracle()
    ctx_d, ctx_h = uv.Context(0), uv.Context(0)
    ctx_h.set_option("host_repack", 1)
    sizes = [1, 2, 3, 7, 16, 17, 33, 100, 1000, 1025, 5000, 40000]
    probs = [.05, .05, .05, .1, .1, .1, .1, .15, .15, .05, .07, .03]
    traced = 0
    for case in range(180):
        kind = KINDS[case % len(KINDS)]
        n = int(rng.choice(sizes, p=probs))
        m = make_mesh(rng, n, kind)
        tag = f"case {case} ({kind}, {n} triangles)"
        td, nd, idd = ctx_d.build_bvh(m)
        th, nh, idh = B.build_bvh(m)
        assert np.array_equal(idd, idh) and nd.tobytes() == nh.tobytes() and td.tobytes() == th.tobytes(), tag + ": trees differ"
        images, errors = [], []
        for ctx in (ctx_h, ctx_d):
            try:
                ctx.upload_scene(th, nh, idh)
                images.append((ctx.scene_info(), ctx.get_option("scene_tame"), ctx.read(uv.BUF.PAIRS).tobytes(), ctx.read(uv.BUF.WTRIS).tobytes()))
                errors.append(None)
            except uv.UvrtError as e:           # e.g. a comb deeper than the traversal stack: both must refuse
                images.append(None)
                errors.append(e.code)
        assert errors[0] == errors[1], tag + f": upload errors differ {errors}"
        if errors[0] is not None:
            continue
        assert images[0] == images[1], tag + ": repacked scenes differ"
        if n > 5000:
            continue
        # rays from a lamp inside the mesh's bounding box
        v = np.concatenate([m[:, 0:3], m[:, 4:7], m[:, 8:11]]).astype(np.float64)
        lo, hi = v.min(axis=0), v.max(axis=0)
        lp = tuple(np.float32(x) for x in (lo + (hi - lo) * rng.uniform(0.3, 0.7, 3)))
        length = np.float32(max(float(hi[1] - lo[1]) * 0.2, 1e-30))
        P = 70000                                # above the binning threshold (65,536 rays)
        seed = int(rng.integers(1 << 32))
        want = np.zeros(P, dtype=T.RAY_DT)
        O.orc_generate(T.ptr(want), 0, P, lp[0], lp[1], lp[2], length, seed, None)
        temp = np.zeros(n, dtype=np.int32)
        O.orc_extend(T.ptr(temp), T.ptr(th), T.ptr(want), T.ptr(nh), T.ptr(idh), P, 0, None)
        for binned in (1, 0):
            ctx_d.set_option("bin_rays", binned)
            ctx_d.reset(False)
            ctx_d.trace_counts(lp, length, 0, P, seed)
            got = ctx_d.read(uv.BUF.RAYS, P)
            assert got.tobytes() == want.tobytes(), tag + f": rays / hits differ (binned={binned})"
            assert np.array_equal(ctx_d.read(uv.BUF.COUNTS), temp), tag + ": counts differ"
        traced += 1
    assert traced > 100
    ctx_d.close()
    ctx_h.close()


def test_upload_of_corrupted_trees_agrees_with_host_validation(uv):
    """Random damage to the node array / triIdx of the room (wild child indices, cycles, leaves that run off the
    end, inner nodes turned into leaves and back): the device-side tree walk must reach the same verdict as the
    host-side validation -- the same error class, or the same repacked scene -- and must always return."""
    sim = uv.Sim(asset_root=T.DATA)
    sim.load_mesh("testroomopt")
    tris, nodes, tri_idx = (a.copy() for a in sim.mesh_data())
    rng = np.random.default_rng(5)
    ctx_d, ctx_h = uv.Context(0), uv.Context(0)
    ctx_h.set_option("host_repack", 1)
    n_nodes, n_tris = len(nodes), tris.shape[0]
    verdicts = {"ok": 0, "error": 0}
    for case in range(120):
        nd, ix = nodes.copy(), tri_idx.copy()
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, n_nodes))
            kind = int(rng.integers(0, 6))
            if kind == 0:
                nd[k]["leftFirst"] = rng.integers(0, 2 ** 32, dtype=np.uint64).astype(np.uint32)
            elif kind == 1:
                nd[k]["leftFirst"] = rng.integers(0, n_nodes)                 # likely a second parent or a cycle
            elif kind == 2:
                nd[k]["triCount"] = rng.integers(0, 5)
            elif kind == 3:
                nd[k]["triCount"] = rng.integers(n_tris, 2 ** 32, dtype=np.uint64).astype(np.uint32)
            elif kind == 4:
                ix[int(rng.integers(0, n_tris))] = rng.integers(n_tris, 2 ** 32, dtype=np.uint64).astype(np.uint32)
            else:
                nd[k]["leftFirst"], nd[k]["triCount"] = 0, 0                  # a zeroed slot: children 0 and 1
        out = []
        for ctx in (ctx_h, ctx_d):
            try:
                ctx.upload_scene(tris, nd, ix)
                out.append(("ok", ctx.scene_info(), ctx.read(uv.BUF.PAIRS).tobytes(), ctx.read(uv.BUF.WTRIS).tobytes()))
            except uv.UvrtError as e:
                out.append(("error", e.code))
        assert out[0][0] == out[1][0], f"case {case}: host says {out[0][:2]}, device says {out[1][:2]}"
        if out[0][0] == "ok":
            assert out[0] == out[1], f"case {case}: repacked scenes differ"
        else:
            assert out[0][1] == out[1][1] == -1
        verdicts[out[0][0]] += 1
    assert verdicts["error"] > 20
    # and both contexts still work
    for ctx in (ctx_h, ctx_d):
        ctx.upload_scene(tris, nodes, tri_idx)
        assert ctx.scene_info()["inner"] == 44708
    ctx_d.close()
    ctx_h.close()
    sim.close()
