"""Shared helpers for the test-suite: ctypes bindings to the CHECKER libraries
(oracle/_build/libuvrt_oracle.so = plain-C port, oracle/_ref/libuvrt_ref.so = the reference's
own sources compiled by oracle/build_ref.sh) and an independent numpy GLB reader.

Nothing here is product code; the product path is small-project-uv-robot-ray-tracer_b200/.
"""
import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
DATA = os.path.join(ROOT, "data")
ROOM = os.path.join(DATA, "rooms", "testroomopt.glb")

RAY_DT = np.dtype([("dir", "<f4", 3), ("orig", "<f4", 3), ("dist", "<f4"), ("triID", "<u4")])
NODE_DT = np.dtype([("min", "<f4", 3), ("leftFirst", "<u4"), ("max", "<f4", 3), ("triCount", "<u4")])
assert RAY_DT.itemsize == 32 and NODE_DT.itemsize == 32


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("innerVisits", C.c_uint64), ("leafVisits", C.c_uint64),
                ("triTests", C.c_uint64), ("hits", C.c_uint64), ("maxStack", C.c_uint32)]


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def build_checkers():
    subprocess.run(["bash", os.path.join(ORACLE_DIR, "build_ref.sh")], check=True,
                   stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "_build", "libuvrt_oracle.so")
        if not os.path.exists(path):
            build_checkers()
        lib = C.CDLL(path)
        lib.orc_wang_hash.restype = C.c_uint32
        lib.orc_wang_hash.argtypes = [C.c_uint32]
        lib.orc_generate.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_uint32, C.POINTER(C.c_uint32)]
        lib.orc_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int, C.POINTER(Counters)]
        lib.orc_brute_force.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        lib.orc_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int]
        lib.orc_compute_dosage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
        lib.orc_dosage_to_color.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
        lib.orc_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_fnv1a64.restype = C.c_uint64
        lib.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        lib.orc_fnv_hits.restype = C.c_uint64
        lib.orc_fnv_hits.argtypes = [C.c_void_p, C.c_int64]
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        _oracle = lib
    return _oracle


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libuvrt_ref.so"))


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libuvrt_ref.so"))
        lib.ref_generate.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_uint, C.POINTER(C.c_uint)]
        lib.ref_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_longlong, C.c_int, C.c_int]
        lib.ref_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int]
        lib.ref_compute_dosage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
        lib.ref_dosage_to_color.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
        lib.ref_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_num_threads.restype = C.c_int
        lib.ref_set_num_threads.argtypes = [C.c_int]
        lib.ref_bvh_build.restype = C.c_int
        lib.ref_bvh_build.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _ref = lib
    return _ref


# ---- independent GLB reader (numpy) -- mirrors what mesh.cpp:28-87 extracts ---------------
def load_glb_tris(path):
    """Returns (N,16) float32: v0.xyz,pad,v1.xyz,pad,v2.xyz,pad,centroid.xyz,pad (pads/centroid zero)."""
    b = open(path, "rb").read()
    magic, version, total = struct.unpack_from("<4sII", b, 0)
    assert magic == b"glTF" and version == 2
    jlen, jtype = struct.unpack_from("<II", b, 12)
    js = json.loads(b[20:20 + jlen])
    off = 20 + jlen
    blen, btype = struct.unpack_from("<II", b, off)
    binchunk = b[off + 8: off + 8 + blen]
    prim = js["meshes"][0]["primitives"][0]

    def view(acc_idx):
        acc = js["accessors"][acc_idx]
        bv = js["bufferViews"][acc["bufferView"]]
        start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
        return acc, start

    pacc, pstart = view(prim["attributes"]["POSITION"])
    pos = np.frombuffer(binchunk, dtype="<f4", count=pacc["count"] * 3, offset=pstart).reshape(-1, 3)
    iacc, istart = view(prim["indices"])
    idt = {5123: "<u2", 5125: "<u4"}[iacc["componentType"]]
    idx = np.frombuffer(binchunk, dtype=idt, count=iacc["count"], offset=istart).astype(np.int64)
    n = iacc["count"] // 3
    tris = np.zeros((n, 16), dtype=np.float32)
    idx = idx[: n * 3].reshape(n, 3)
    tris[:, 0:3] = pos[idx[:, 0]]
    tris[:, 4:7] = pos[idx[:, 1]]
    tris[:, 8:11] = pos[idx[:, 2]]
    return tris


def floor_height(tris):
    """mesh.cpp:100-136 restated with numpy (fp32 arithmetic as in the source)."""
    ys = tris[:, [1, 5, 9]].reshape(-1).astype(np.float32)
    f = np.float32
    minv = min(f(0.0), ys.min())
    rng = f(0.0) - minv
    bins = 48
    hist = np.zeros(bins, dtype=np.int64)
    for j in range(bins):
        lo = f(f(j) * rng) / f(bins) + minv
        hi = f(f(j + 1) * rng) / f(bins) + minv
        hist[j] = np.count_nonzero((lo < ys) & (ys < hi))
    mi = int(np.argmax(hist)) if hist.max() > 0 else -1
    return f(f(f(mi) + f(0.5)) * rng / f(bins) + minv)


def ref_build_bvh(tris):
    """Runs the reference's builder. Returns (tris_with_centroids, nodes[NODE_DT], triIdx)."""
    t = np.ascontiguousarray(tris.copy())
    # the builder needs 64-byte aligned triangles for its SSE loads
    raw = np.zeros(t.nbytes + 64, dtype=np.uint8)
    o = (-raw.ctypes.data) % 64
    ta = raw[o:o + t.nbytes].view(np.float32).reshape(t.shape)
    ta[:] = t
    n = t.shape[0]
    cap = 2 * n + 64
    nodes = np.zeros(cap, dtype=NODE_DT)
    tri_idx = np.zeros(n, dtype=np.uint32)
    used = ref().ref_bvh_build(ptr(ta), n, ptr(nodes), cap, ptr(tri_idx))
    return ta.copy(), nodes[:used], tri_idx


def fnv(a):
    a = np.ascontiguousarray(a)
    return int(oracle().orc_fnv1a64(ptr(a), a.nbytes))


def reachable_preorder(nodes):
    """Indices of nodes reachable from the root, pre-order, left child first."""
    out = []
    stack = [0]
    lf = nodes["leftFirst"]
    tc = nodes["triCount"]
    while stack:
        i = stack.pop()
        out.append(i)
        if tc[i] == 0:
            stack.append(int(lf[i]) + 1)
            stack.append(int(lf[i]))
    return np.array(out, dtype=np.int64)
