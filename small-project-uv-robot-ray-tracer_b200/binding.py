"""ctypes bindings for include/uvrt.h and include/uvrt_host.h."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

RAY_DTYPE = np.dtype([("dir", "<f4", 3), ("orig", "<f4", 3), ("dist", "<f4"), ("triID", "<u4")])
NODE_DTYPE = np.dtype([("min", "<f4", 3), ("leftFirst", "<u4"), ("max", "<f4", 3), ("triCount", "<u4")])


class UvrtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"uvrt error {code}: {message}")
        self.code = code


class BUF:
    RAYS, COUNTS, SUM, MAX, DOSE, COLOR, PAIRS, WTRIS, MATRIX = range(9)


class STAGE:
    GENERATE, EXTEND, ACCUMULATE, SHADE, COLOR, RESET, BIN = range(7)


def build_dir():
    return os.path.join(_HERE, "_build")


def include_dir():
    return os.path.join(_ROOT, "include")


def build(verbose=False):
    """Compiles libuvrt.so (nvcc, sm_100a), libuvrt_host.so and uvrt_cli in-tree."""
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building the uvrt libraries failed")


def declared_symbols(header):
    """Names of the functions a header under include/ declares."""
    text = open(os.path.join(include_dir(), header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(uvrt_[a-z0-9_]+)\s*\(", text)))


_lib = None
_host = None


def _load(name):
    path = os.path.join(build_dir(), name)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `make -C {_HERE}` (or __graft_entry__.build()) first; "
                           "there is no fallback implementation")
    return C.CDLL(path)  # RTLD_LOCAL: the C++ class names also exist in the checker libraries


class SimParams(C.Structure):
    _fields_ = [("photonCount", C.c_int), ("maxIterations", C.c_int), ("lightIntensity", C.c_float),
                ("minDosage", C.c_float), ("minPower", C.c_float), ("lightLength", C.c_float),
                ("lightHeight", C.c_float), ("viewMode", C.c_int), ("thresholdView", C.c_int),
                ("photonsPerLight", C.c_int), ("currIterations", C.c_int), ("photonMapSize", C.c_int),
                ("seedState", C.c_uint32), ("finishedComputation", C.c_int)]


def lib():
    """libuvrt.so with argument types declared."""
    global _lib
    if _lib is not None:
        return _lib
    L = _load("libuvrt.so")
    vp, i, f, i64, u32 = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint32
    sig = {
        "uvrt_device_count": (i, [C.POINTER(i)]),
        "uvrt_create": (i, [C.POINTER(vp), i]),
        "uvrt_destroy": (None, [vp]),
        "uvrt_last_error": (C.c_char_p, [vp]),
        "uvrt_device_info": (i, [vp, C.c_char_p, C.c_size_t, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "uvrt_upload_scene": (i, [vp, vp, i, vp, i, vp]),
        "uvrt_build_bvh": (i, [vp, vp, i, vp, i, vp, C.POINTER(u32), vp]),
        "uvrt_reset": (i, [vp, i]),
        "uvrt_generate": (i, [vp, f, f, f, f, i64, i64, u32]),
        "uvrt_extend": (i, [vp, i64]),
        "uvrt_accumulate": (i, [vp, f]),
        "uvrt_trace": (i, [vp, f, f, f, f, f, i64, i64, u32]),
        "uvrt_trace_counts": (i, [vp, f, f, f, f, i64, i64, u32]),
        "uvrt_seed_chain": (i, [vp, vp, i, f, u32, vp]),
        "uvrt_shade": (i, [vp, i, i, f]),
        "uvrt_color": (i, [vp, f, i]),
        "uvrt_read": (i, [vp, i, vp, C.c_size_t]),
        "uvrt_write": (i, [vp, i, vp, C.c_size_t]),
        "uvrt_sync": (i, [vp]),
        "uvrt_comm_unique_id": (i, [vp]),
        "uvrt_comm_init": (i, [vp, vp, i, i]),
        "uvrt_reduce": (i, [vp]),
        "uvrt_reduce_counts": (i, [vp]),
        "uvrt_probe_cost": (i, [vp, f, f, f, f, u32, i, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "uvrt_matrix_reserve": (i, [vp, i]),
        "uvrt_matrix_begin": (i, [vp, i]),
        "uvrt_trace_row": (i, [vp, i, f, f, f, f, i64, i64, u32]),
        "uvrt_matrix_fold": (i, [vp, vp, i, i]),
        "uvrt_timeline_dump": (i, [vp, C.c_char_p]),
        "uvrt_set_option": (i, [vp, C.c_char_p, i]),
        "uvrt_get_option": (i, [vp, C.c_char_p, C.POINTER(i)]),
        "uvrt_stage_time": (i, [vp, i, C.POINTER(C.c_double), C.POINTER(i64)]),
        "uvrt_stage_time_reset": (i, [vp]),
        "uvrt_launch_count": (i64, [vp]),
        "uvrt_mark": (i, [vp, i]),
        "uvrt_elapsed_ms": (i, [vp, i, i, C.POINTER(f)]),
        "uvrt_scene_info": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "uvrt_selftest_division": (i, [vp, i, i, vp]),
        "uvrt_fast_stats": (i, [vp, vp, i]),
        "uvrt_flush_l2": (i, [vp]),
        "uvrt_scene_upload_bytes": (i64, [vp]),
        "uvrt_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def host():
    """libuvrt_host.so with argument types declared."""
    global _host
    if _host is not None:
        return _host
    lib()  # libuvrt_host.so links against libuvrt.so
    H = _load("libuvrt_host.so")
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "uvrt_sim_create": (i, [C.POINTER(vp), C.c_char_p, i]),
        "uvrt_sim_destroy": (None, [vp]),
        "uvrt_sim_last_error": (C.c_char_p, [vp]),
        "uvrt_sim_load_mesh": (i, [vp, C.c_char_p]),
        "uvrt_sim_set_triangles": (i, [vp, vp, i]),
        "uvrt_sim_set_device_bvh": (i, [vp, i]),
        "uvrt_sim_set_whole_scene": (i, [vp, i]),
        "uvrt_sim_mesh_info": (i, [vp, C.POINTER(i), C.POINTER(f), C.POINTER(C.c_uint)]),
        "uvrt_sim_mesh_data": (i, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "uvrt_sim_load_route": (i, [vp, C.c_char_p]),
        "uvrt_sim_save_route": (i, [vp, C.c_char_p]),
        "uvrt_sim_get_params": (i, [vp, C.POINTER(SimParams)]),
        "uvrt_sim_set_params": (i, [vp, C.POINTER(SimParams)]),
        "uvrt_sim_get_positions": (i, [vp, vp, i, C.POINTER(i)]),
        "uvrt_sim_set_positions": (i, [vp, vp, i]),
        "uvrt_sim_init": (i, [vp, C.c_char_p]),
        "uvrt_sim_reset_dosage_map": (i, [vp]),
        "uvrt_sim_compute_dosage_map": (i, [vp]),
        "uvrt_sim_compute_single": (i, [vp, f, f, f, i, i]),
        "uvrt_sim_shade": (i, [vp]),
        "uvrt_sim_tick": (i, [vp, C.POINTER(i)]),
        "uvrt_sim_run": (i, [vp, vp, i]),
        "uvrt_sim_calibrate": (i, [vp, f, f, f, C.POINTER(f)]),
        "uvrt_sim_read_dose": (i, [vp, vp, i]),
        "uvrt_sim_save_dosage_map": (i, [vp, C.c_char_p]),
        "uvrt_sim_save_checkpoint": (i, [vp, C.c_char_p]),
        "uvrt_sim_load_checkpoint": (i, [vp, C.c_char_p]),
        "uvrt_sim_set_shard": (i, [vp, i, i]),
        "uvrt_sim_set_shard_parts": (i, [vp, i]),
        "uvrt_sim_shard_parts": (i, [vp]),
        "uvrt_sim_set_cost_aware": (i, [vp, i]),
        "uvrt_host_plan_shards": (i, [vp, i, i, vp]),
        "uvrt_sim_set_seed": (i, [vp, C.c_uint32]),
        "uvrt_host_seed_after_launch": (C.c_uint32, [f, f, f, f, C.c_uint32]),
        "uvrt_host_shard_owner": (i, [C.c_longlong, i, i]),
        "uvrt_sim_reduce": (i, [vp]),
        "uvrt_sim_ctx": (vp, [vp]),
        "uvrt_sim_rays_traced": (C.c_int64, [vp]),
        "uvrt_host_build_bvh": (i, [vp, i, vp, i, vp, C.POINTER(C.c_uint)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(H, name)
        fn.restype = res
        fn.argtypes = args
    _host = H
    return H


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """A uvrt_ctx (one CUDA device, one stream)."""

    def __init__(self, device=0, handle=None):
        self.L = lib()
        self.owned = handle is None
        if handle is None:
            h = C.c_void_p()
            rc = self.L.uvrt_create(C.byref(h), device)
            if rc != 0:
                raise UvrtError(rc, self.L.uvrt_last_error(None).decode())
            handle = h
        self.h = handle
        self.n_tris = 0

    def close(self):
        if self.owned and self.h:
            self.L.uvrt_destroy(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        if rc != 0:
            raise UvrtError(rc, self.L.uvrt_last_error(self.h).decode())

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        self.check(self.L.uvrt_device_info(self.h, name, 256, C.byref(sm), C.byref(ma), C.byref(mi)))
        return {"name": name.value.decode(), "sms": sm.value, "cc": (ma.value, mi.value)}

    def upload_scene(self, tris, nodes, tri_idx):
        tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 16)
        nodes = np.ascontiguousarray(nodes)
        tri_idx = np.ascontiguousarray(tri_idx, dtype=np.uint32)
        assert nodes.dtype.itemsize == 32 or nodes.dtype == np.uint8
        n_nodes = nodes.nbytes // 32
        self.check(self.L.uvrt_upload_scene(self.h, _p(tris), tris.shape[0], _p(nodes), n_nodes, _p(tri_idx)))
        self.n_tris = tris.shape[0]

    def build_bvh(self, tris, out=None):
        """Device BVH build. Returns (tris with centroids, nodes, triIdx) like binding.build_bvh().
        out: optional preallocated (tris (n,16) f32, nodes (2n+64) NODE_DTYPE, triIdx (n) u32) to write into."""
        t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 16)
        n = t.shape[0]
        cap = 2 * n + 64
        if out is None:
            out = (np.zeros_like(t), np.zeros(cap, dtype=NODE_DTYPE), np.zeros(n, dtype=np.uint32))
        out_tris, nodes, tri_idx = out
        assert out_tris.shape == t.shape and len(nodes) >= cap and len(tri_idx) == n
        used = C.c_uint32()
        self.check(self.L.uvrt_build_bvh(self.h, _p(t), n, _p(nodes), cap, _p(tri_idx), C.byref(used), _p(out_tris)))
        return out_tris, nodes[: used.value], tri_idx

    def scene_info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.check(self.L.uvrt_scene_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"inner": a.value, "leaves": b.value, "depth": c.value, "stack": d.value}

    def reset(self, reset_color=True):
        self.check(self.L.uvrt_reset(self.h, int(reset_color)))

    def generate(self, lp, light_length, first_ray, n_rays, seed_in):
        self.check(self.L.uvrt_generate(self.h, lp[0], lp[1], lp[2], light_length, first_ray, n_rays, seed_in))

    def extend(self, n_rays):
        self.check(self.L.uvrt_extend(self.h, n_rays))

    def accumulate(self, duration):
        self.check(self.L.uvrt_accumulate(self.h, duration))

    def trace(self, lp, light_length, duration, first_ray, n_rays, seed_in):
        self.check(self.L.uvrt_trace(self.h, lp[0], lp[1], lp[2], light_length, duration, first_ray, n_rays, seed_in))

    def trace_counts(self, lp, light_length, first_ray, n_rays, seed_in):
        self.check(self.L.uvrt_trace_counts(self.h, lp[0], lp[1], lp[2], light_length, first_ray, n_rays, seed_in))

    def seed_chain(self, positions, light_length, seed_in):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(pos.shape[0] + 1, dtype=np.uint32)
        self.check(self.L.uvrt_seed_chain(self.h, _p(pos), pos.shape[0], light_length, seed_in, _p(out)))
        return out

    def shade(self, use_max, photons_per_light, scaled_power):
        self.check(self.L.uvrt_shade(self.h, int(use_max), photons_per_light, scaled_power))

    def color(self, min_value, threshold_view):
        self.check(self.L.uvrt_color(self.h, min_value, int(threshold_view)))

    def read(self, what, n_rays=None):
        n = self.n_tris
        if what == BUF.RAYS:
            out = np.zeros(n_rays, dtype=RAY_DTYPE)
        elif what == BUF.COUNTS:
            out = np.zeros(n, dtype=np.int32)
        elif what in (BUF.SUM, BUF.MAX):
            out = np.zeros(n, dtype=np.float64)
        elif what == BUF.DOSE:
            out = np.zeros(n, dtype=np.float32)
        elif what == BUF.PAIRS:
            out = np.zeros((max(self.scene_info()["inner"], 1), 16), dtype=np.uint32)
        elif what == BUF.WTRIS:
            out = np.zeros((n_rays if n_rays is not None else n, 16), dtype=np.uint32)
        elif what == BUF.MATRIX:
            out = np.zeros((n_rays, n), dtype=np.int32)        # n_rays = rows
        else:
            out = np.zeros((n, 9), dtype=np.float32)
        self.check(self.L.uvrt_read(self.h, what, _p(out), out.nbytes))
        return out

    def write(self, what, arr):
        arr = np.ascontiguousarray(arr)
        self.check(self.L.uvrt_write(self.h, what, _p(arr), arr.nbytes))

    def sync(self):
        self.check(self.L.uvrt_sync(self.h))

    def set_option(self, key, value):
        self.check(self.L.uvrt_set_option(self.h, key.encode(), int(value)))

    def get_option(self, key):
        v = C.c_int()
        self.check(self.L.uvrt_get_option(self.h, key.encode(), C.byref(v)))
        return v.value

    def stage_time(self, stage):
        ms, n = C.c_double(), C.c_int64()
        self.check(self.L.uvrt_stage_time(self.h, stage, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stage_time_reset(self):
        self.check(self.L.uvrt_stage_time_reset(self.h))

    def flush_l2(self):
        self.check(self.L.uvrt_flush_l2(self.h))

    def scene_upload_bytes(self):
        return int(self.L.uvrt_scene_upload_bytes(self.h))

    def launch_count(self):
        return int(self.L.uvrt_launch_count(self.h))

    def mark(self, slot):
        self.check(self.L.uvrt_mark(self.h, slot))

    def elapsed_ms(self, a, b):
        ms = C.c_float()
        self.check(self.L.uvrt_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def selftest_division(self, blocks=148 * 8, iters=4096):
        out = np.zeros(3, dtype=np.uint64)
        self.check(self.L.uvrt_selftest_division(self.h, blocks, iters, _p(out)))
        return [int(x) for x in out]

    def fast_stats(self, reset=False):
        """{'cert_fallbacks', 'ineligible', 'check_mismatches'} of the certified fast extend since the last reset."""
        out = np.zeros(3, dtype=np.uint64)
        self.check(self.L.uvrt_fast_stats(self.h, _p(out), int(reset)))
        return {"cert_fallbacks": int(out[0]), "ineligible": int(out[1]), "check_mismatches": int(out[2])}

    def comm_init(self, id128, rank, n_ranks):
        buf = (C.c_char * 128).from_buffer_copy(bytes(id128))
        self.check(self.L.uvrt_comm_init(self.h, buf, rank, n_ranks))

    def reduce(self):
        self.check(self.L.uvrt_reduce(self.h))

    def reduce_counts(self):
        self.check(self.L.uvrt_reduce_counts(self.h))

    def probe_cost(self, lp, light_length, seed_in=0, n_rays=8192):
        a, b = C.c_double(), C.c_double()
        self.check(self.L.uvrt_probe_cost(self.h, lp[0], lp[1], lp[2], light_length, seed_in, n_rays, C.byref(a), C.byref(b)))
        return a.value, b.value

    def matrix_begin(self, rows):
        self.check(self.L.uvrt_matrix_begin(self.h, rows))

    def trace_row(self, row, lp, light_length, first_ray, n_rays, seed_in):
        self.check(self.L.uvrt_trace_row(self.h, row, lp[0], lp[1], lp[2], light_length, first_ray, n_rays, seed_in))

    def matrix_fold(self, durations, reduce=False):
        d = np.ascontiguousarray(durations, dtype=np.float32)
        self.check(self.L.uvrt_matrix_fold(self.h, _p(d), d.shape[0], int(reduce)))

    def timeline_dump(self, path):
        self.check(self.L.uvrt_timeline_dump(self.h, str(path).encode()))


def comm_unique_id():
    buf = (C.c_char * 128)()
    rc = lib().uvrt_comm_unique_id(buf)
    if rc != 0:
        raise UvrtError(rc, lib().uvrt_last_error(None).decode())
    return bytes(buf)


class Sim:
    """A uvrt_sim: the reference's Mesh + RayTracer pair (host drop-in classes)."""

    def __init__(self, asset_root=None, device=0):
        self.H = host()
        h = C.c_void_p()
        rc = self.H.uvrt_sim_create(C.byref(h), asset_root.encode() if asset_root else None, device)
        if rc != 0:
            raise UvrtError(rc, "uvrt_sim_create failed")
        self.h = h

    def close(self):
        if self.h:
            self.H.uvrt_sim_destroy(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        if rc != 0:
            raise UvrtError(rc, self.H.uvrt_sim_last_error(self.h).decode())

    def load_mesh(self, model_file):
        self.check(self.H.uvrt_sim_load_mesh(self.h, model_file.encode()))

    def set_whole_scene(self, on=True):
        self.check(self.H.uvrt_sim_set_whole_scene(self.h, int(on)))

    def set_device_bvh(self, on=True):
        self.check(self.H.uvrt_sim_set_device_bvh(self.h, int(on)))

    def set_triangles(self, tris):
        tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 16)
        self.check(self.H.uvrt_sim_set_triangles(self.h, _p(tris), tris.shape[0]))

    def mesh_info(self):
        n, fh, used = C.c_int(), C.c_float(), C.c_uint()
        self.check(self.H.uvrt_sim_mesh_info(self.h, C.byref(n), C.byref(fh), C.byref(used)))
        return {"triangles": n.value, "floor": np.float32(fh.value), "nodesUsed": used.value}

    def mesh_data(self):
        info = self.mesh_info()
        t, nd, ti = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.check(self.H.uvrt_sim_mesh_data(self.h, C.byref(t), C.byref(nd), C.byref(ti)))
        n = info["triangles"]
        tris = np.ctypeslib.as_array(C.cast(t, C.POINTER(C.c_float)), shape=(n, 16)).copy()
        nodes = np.frombuffer(C.string_at(nd, info["nodesUsed"] * 32), dtype=NODE_DTYPE).copy()
        tri_idx = np.ctypeslib.as_array(C.cast(ti, C.POINTER(C.c_uint32)), shape=(n,)).copy()
        return tris, nodes, tri_idx

    def load_route(self, name):
        self.check(self.H.uvrt_sim_load_route(self.h, name.encode()))

    def save_route(self, name):
        self.check(self.H.uvrt_sim_save_route(self.h, name.encode()))

    @property
    def params(self):
        p = SimParams()
        self.check(self.H.uvrt_sim_get_params(self.h, C.byref(p)))
        return p

    def set_params(self, **kw):
        p = self.params
        for k, v in kw.items():
            setattr(p, k, v)
        self.check(self.H.uvrt_sim_set_params(self.h, C.byref(p)))

    @property
    def positions(self):
        n = C.c_int()
        self.check(self.H.uvrt_sim_get_positions(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 3), dtype=np.float32)
        self.check(self.H.uvrt_sim_get_positions(self.h, _p(out), n.value, C.byref(n)))
        return out

    def set_positions(self, xyd):
        xyd = np.ascontiguousarray(xyd, dtype=np.float32).reshape(-1, 3)
        self.check(self.H.uvrt_sim_set_positions(self.h, _p(xyd), xyd.shape[0]))

    def init(self, route_name=None):
        self.check(self.H.uvrt_sim_init(self.h, route_name.encode() if route_name else None))

    @property
    def ctx(self):
        c = Context(handle=C.c_void_p(self.H.uvrt_sim_ctx(self.h)))
        c.n_tris = self.mesh_info()["triangles"]
        return c

    def reset_dosage_map(self):
        self.check(self.H.uvrt_sim_reset_dosage_map(self.h))

    def compute_dosage_map(self):
        self.check(self.H.uvrt_sim_compute_dosage_map(self.h))

    def compute_single(self, x, y, duration, photons, triangle_count):
        self.check(self.H.uvrt_sim_compute_single(self.h, x, y, duration, photons, triangle_count))

    def shade(self):
        self.check(self.H.uvrt_sim_shade(self.h))

    def tick(self):
        fin = C.c_int()
        self.check(self.H.uvrt_sim_tick(self.h, C.byref(fin)))
        return bool(fin.value)

    def run(self, dose=None):
        n = self.mesh_info()["triangles"]
        if dose is None:
            dose = np.zeros(n, dtype=np.float32)
        self.check(self.H.uvrt_sim_run(self.h, _p(dose), dose.shape[0]))
        return dose

    def calibrate(self, measure_power, measure_height, measure_dist):
        out = C.c_float()
        self.check(self.H.uvrt_sim_calibrate(self.h, measure_power, measure_height, measure_dist, C.byref(out)))
        return out.value

    def save_dosage_map(self, base_path):
        self.check(self.H.uvrt_sim_save_dosage_map(self.h, str(base_path).encode()))

    def save_checkpoint(self, path):
        self.check(self.H.uvrt_sim_save_checkpoint(self.h, str(path).encode()))

    def load_checkpoint(self, path):
        self.check(self.H.uvrt_sim_load_checkpoint(self.h, str(path).encode()))

    def read_dose(self):
        n = self.mesh_info()["triangles"]
        out = np.zeros(n, dtype=np.float32)
        self.check(self.H.uvrt_sim_read_dose(self.h, _p(out), n))
        return out

    def set_shard(self, rank, count):
        self.check(self.H.uvrt_sim_set_shard(self.h, rank, count))

    def set_shard_parts(self, parts):
        self.check(self.H.uvrt_sim_set_shard_parts(self.h, parts))

    def set_cost_aware(self, on=True):
        self.check(self.H.uvrt_sim_set_cost_aware(self.h, int(on)))

    def shard_parts(self):
        return int(self.H.uvrt_sim_shard_parts(self.h))

    def set_seed(self, seed):
        self.check(self.H.uvrt_sim_set_seed(self.h, seed))

    def reduce(self):
        self.check(self.H.uvrt_sim_reduce(self.h))

    def rays_traced(self):
        return int(self.H.uvrt_sim_rays_traced(self.h))


def build_bvh(tris):
    """The host builder on its own. Returns (tris with centroids, nodes, triIdx)."""
    t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 16).copy()
    n = t.shape[0]
    cap = 2 * n + 64
    nodes = np.zeros(cap, dtype=NODE_DTYPE)
    tri_idx = np.zeros(n, dtype=np.uint32)
    used = C.c_uint()
    rc = host().uvrt_host_build_bvh(_p(t), n, _p(nodes), cap, _p(tri_idx), C.byref(used))
    if rc != 0:
        raise UvrtError(rc, "uvrt_host_build_bvh failed")
    return t, nodes[: used.value], tri_idx
