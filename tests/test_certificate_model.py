"""The certified fast extend's argument, checked on the host (no GPU): tools/traversal_lab.c restates the scheme of
csrc/uvrt_fast.cuh in plain C -- conservative 15-bit boxes, culling with a margin, the reference's exact triangle test
and exact leaf-box test, the near-tie certificate -- and compares every ray with the reference-order traversal
(extend.cl:40-81 in the reference's arithmetic).  The claim: a CERTIFIED ray never differs; uncertified rays are rare.
The GPU suite checks the same on the kernel itself ("fast_check"); this is the CPU-side counterpart.  (The host model
pads its quantised boxes by one cell where the kernel pads by two, and applies the exact leaf-box test when a hit is
accepted where the kernel applies it to the winner after the loop: both are instances of the same argument.)"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import pytest

import uvrt_testlib as T

sys.path.insert(0, os.path.join(T.ROOT, "tools"))
import traversal_lab as TL  # noqa: E402

REL, ABS = 2.0 ** -12, 2.0 ** -14      # kFastRel, kFastAbs of csrc/uvrt_fast.cuh


def _scene(kind, uv):
    f32 = np.float32
    if kind == "room":
        sim = uv.Sim(asset_root=T.DATA)
        sim.load_mesh("testroomopt")
        sim.load_route("route")
        tris, nodes, tri_idx = sim.mesh_data()
        floor = sim.mesh_info()["floor"]
        p = sim.params
        lamps = [(f32(x), f32(f32(floor) + f32(p.lightHeight)), f32(y)) for x, y, _ in sim.positions]
        length = p.lightLength
        sim.close()
        return tris, nodes, tri_idx, lamps, length
    from soup import make_soup, soup_route
    B = importlib.import_module("small-project-uv-robot-ray-tracer_b200.binding")
    tris, nodes, tri_idx = B.build_bvh(make_soup(200_000))
    return tris, nodes, tri_idx, [(f32(x), f32(0.5), f32(z)) for x, z, _ in soup_route()], 1.0


@pytest.mark.parametrize("kind,positions,n_rays", [("room", (0, 11), 150_000), ("soup200k", (5,), 100_000)])
def test_certified_rays_equal_the_reference_order_traversal(uv, checkers, kind, positions, n_rays):
    tris, nodes, tri_idx, lamps, length = _scene(kind, uv)
    L, O = TL.lab(), T.oracle()
    for k in positions:
        lp = lamps[k]
        rays = np.zeros(n_rays, dtype=T.RAY_DT)
        O.orc_generate(T.ptr(rays), 0, n_rays, lp[0], lp[1], lp[2], np.float32(length), 7 * k, None)
        ref = rays.copy()
        s = TL.Stats()
        L.lab_exact_stats(T.ptr(tris), T.ptr(ref), T.ptr(nodes), T.ptr(tri_idx), n_rays, 0, C.byref(s))
        # the scheme's one numerical assumption: an accepted t is not below its leaf box's exact entry distance by more
        # than REL * t + ABS; observed inversions stay two orders of magnitude inside that
        assert s.accepted > 0
        assert s.maxInvAbs < ABS / 50 and s.maxInvRel < REL / 50
        # the lab's reference-order traversal is the oracle's
        chk = rays.copy()
        temp = np.zeros(len(tris), dtype=np.int32)
        O.orc_extend(T.ptr(temp), T.ptr(tris), T.ptr(chk), T.ptr(nodes), T.ptr(tri_idx), n_rays, 1, None)
        assert chk.tobytes() == ref.tobytes()
        for quant, pop_check in ((1, 0), (1, 1), (0, 1)):
            out = rays.copy()
            f = TL.Stats()
            L.lab_fast(T.ptr(tris), T.ptr(out), T.ptr(ref), T.ptr(nodes), len(nodes), T.ptr(tri_idx), n_rays, quant,
                       REL, ABS, pop_check, C.byref(f))
            assert f.mismatch == 0, f"{f.mismatch} certified rays differ from the reference (quant={quant})"
            assert f.certFail <= n_rays // 1000, "the certificate should fail for a few rays in 10^5 only"
            # an uncertified ray is re-traced in reference order by the kernel, so what the fast traversal alone gets
            # wrong must be covered by certificate failures
            assert f.rawMismatch <= f.certFail
