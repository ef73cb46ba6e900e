"""ctypes binding of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (vk-renderer_b200/) never does.

The oracle restates Assets/Tracer.comp and Assets/Raytracer.comp on the CPU
(oracle/vkrt_oracle.cpp).  The reference has no tests or golden vectors of its own; the pins are the
outputs of its compiled shaders (Assets/Compiled/*.spv) executed by oracle/spirv_interp.py, committed as
tests/golden/spirv_vectors.npz and checked in tests/test_spirv_pins.py; see vkrt_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

WHITTED, PATH = 0, 1
LITERAL, S_LINEAR, S_BVH = 0, 1, 2
SCENE_TRACER, SCENE_RAYTRACER = 0, 1


class Sphere(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("cz", C.c_float), ("r", C.c_float)]


class Plane(C.Structure):
    _fields_ = [("nx", C.c_float), ("ny", C.c_float), ("nz", C.c_float), ("len", C.c_float)]


class Params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("max_depth", C.c_uint32),
                ("integrator", C.c_uint32), ("sphere_mode", C.c_uint32), ("seed", C.c_uint64),
                ("frame_index", C.c_uint32), ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32),
                ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32),
                ("n_threads", C.c_uint32), ("accumulate", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("node_visits", C.c_uint64),
                ("leaf_tests", C.c_uint64), ("paths", C.c_uint64), ("literal_vs_s_mismatch", C.c_uint64)]


def build(fast=False):
    """Compiles the oracle (and oracle/_ref when /root/reference is present). Idempotent."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def _load(fast):
    name = "libvkrt_oracle_fast.so" if fast else "libvkrt_oracle.so"
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    f32p, u32p, u8p, vp = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_void_p
    lib.orc_scene_create.restype = vp
    lib.orc_scene_destroy.argtypes = [vp]
    lib.orc_scene_set_materials.argtypes = [vp, vp, C.c_uint32]
    lib.orc_scene_set_spheres.argtypes = [vp, vp, vp, C.c_uint32]
    lib.orc_scene_set_planes.argtypes = [vp, vp, vp, C.c_uint32]
    lib.orc_scene_set_triangles.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    lib.orc_scene_set_triangle_materials.argtypes = [vp, vp, C.c_uint32]
    lib.orc_scene_use_default.argtypes = [vp, C.c_uint32]
    lib.orc_scene_build_bvh.argtypes = [vp]
    lib.orc_scene_set_bvh.argtypes = [vp, vp, C.c_uint32]
    lib.orc_scene_bvh_nodes.argtypes = [vp]
    lib.orc_scene_bvh_nodes.restype = C.c_uint32
    lib.orc_scene_read_bvh.argtypes = [vp, vp, C.c_size_t]
    lib.orc_render.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp, C.POINTER(Counters)]
    lib.orc_resolve.argtypes = [C.POINTER(Params), vp, vp, vp]
    lib.orc_present.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_uint32]
    for fn in ("orc_sin", "orc_cos", "orc_exp2", "orc_log2"):
        getattr(lib, fn).argtypes = [C.c_float]
        getattr(lib, fn).restype = C.c_float
    lib.orc_pow.argtypes = [C.c_float, C.c_float]
    lib.orc_pow.restype = C.c_float
    lib.orc_pcg_hash.argtypes = [C.c_uint32]
    lib.orc_pcg_hash.restype = C.c_uint32
    lib.orc_frame_key.argtypes = [C.c_uint64, C.c_float, C.c_uint32]
    lib.orc_frame_key.restype = C.c_uint32
    lib.orc_rand_u01.argtypes = [C.c_uint32] * 4
    lib.orc_rand_u01.restype = C.c_float
    lib.orc_sphere_intersect.argtypes = [f32p, f32p, vp]
    lib.orc_sphere_intersect.restype = C.c_float
    lib.orc_plane_intersect_tracer.argtypes = [f32p, f32p, vp]
    lib.orc_plane_intersect_tracer.restype = C.c_float
    lib.orc_plane_intersect_raytracer.argtypes = [f32p, f32p, vp]
    lib.orc_plane_intersect_raytracer.restype = C.c_float
    lib.orc_tri_intersect.argtypes = [f32p, f32p, vp, C.c_float]
    lib.orc_tri_intersect.restype = C.c_float
    lib.orc_primary_ray.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p]
    lib.orc_slab.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p]
    lib.orc_query_spheres.argtypes = [vp, C.c_uint32, f32p, f32p, C.c_float, C.c_float, f32p]
    return lib


_LIBS = {}


def lib(fast=False):
    if fast not in _LIBS:
        _LIBS[fast] = _load(fast)
    return _LIBS[fast]


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """An oracle scene.  Arrays use the same byte layouts as include/vkrt.h."""

    def __init__(self, fast=False):
        self.l = lib(fast)
        self.h = self.l.orc_scene_create()

    def __del__(self):
        try:
            self.l.orc_scene_destroy(self.h)
        except Exception:
            pass

    def use_default(self, which):
        assert self.l.orc_scene_use_default(self.h, which) == 0
        return self

    def set_materials(self, mats):  # (n, 12) float32 with the type as uint32 bits in column 8
        mats = np.ascontiguousarray(mats)
        assert mats.dtype.itemsize * mats.shape[1] == 48
        assert self.l.orc_scene_set_materials(self.h, _ptr(mats), mats.shape[0]) == 0

    def set_spheres(self, spheres, mat_id):
        spheres = np.ascontiguousarray(spheres, dtype=np.float32)
        mat_id = np.ascontiguousarray(mat_id, dtype=np.uint32)
        assert self.l.orc_scene_set_spheres(self.h, _ptr(spheres), _ptr(mat_id), spheres.shape[0]) == 0

    def set_planes(self, planes, mat_id):
        planes = np.ascontiguousarray(planes, dtype=np.float32)
        mat_id = np.ascontiguousarray(mat_id, dtype=np.uint32)
        assert self.l.orc_scene_set_planes(self.h, _ptr(planes), _ptr(mat_id), planes.shape[0]) == 0

    def set_triangles(self, tris, mat_id):  # (n, 12) float32 (3 x vec3 padded to 16 B)
        tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 12)
        assert self.l.orc_scene_set_triangles(self.h, _ptr(tris), tris.shape[0], mat_id) == 0

    def set_triangle_materials(self, mat_ids):  # one material id per triangle (None / empty: all use set_triangles' mat_id)
        m = np.ascontiguousarray(mat_ids if mat_ids is not None else [], dtype=np.uint32)
        assert self.l.orc_scene_set_triangle_materials(self.h, _ptr(m), m.shape[0]) == 0

    def build_bvh(self):
        assert self.l.orc_scene_build_bvh(self.h) == 0
        return self

    def set_bvh(self, nodes):
        """Test hook: walk this hierarchy ((n, 16) float32, bvh_nodes' layout, root = node 0) instead of the oracle's own LBVH."""
        a = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1, 16)
        assert self.l.orc_scene_set_bvh(self.h, _ptr(a), a.shape[0]) == 0
        return self

    def bvh_nodes(self):
        n = self.l.orc_scene_bvh_nodes(self.h)
        out = np.zeros((n, 16), dtype=np.float32)
        if n:
            assert self.l.orc_scene_read_bvh(self.h, _ptr(out), out.nbytes) == 0
        return out

    def query_spheres(self, mode, o, d, bound, eps=1e-3):
        t = C.c_float(0)
        i = self.l.orc_query_spheres(self.h, mode, _f3(o), _f3(d), bound, eps, C.byref(t))
        return i, t.value

    def render(self, frame_data, width, height, spp=4, max_depth=4, integrator=PATH, sphere_mode=LITERAL,
               seed=0, frame_index=0, samples=None, rect=None, n_threads=0, accum=None, want_ids=True,
               want_rgba=True):
        """Returns (accum[h,w,4] f32, hit_ids[h,w] u32 | None, rgba8[h,w,4] u8 | None, Counters)."""
        fd = np.frombuffer(bytes(frame_data), dtype=np.uint8).copy()
        assert fd.nbytes == 96
        p = Params(width=width, height=height, spp=spp, max_depth=max_depth, integrator=integrator,
                   sphere_mode=sphere_mode, seed=seed, frame_index=frame_index, n_threads=n_threads)
        if samples is not None:
            p.sample_begin, p.sample_end = samples
        if rect is not None:
            p.x0, p.y0, p.x1, p.y1 = rect
        if accum is None:
            accum = np.zeros((height, width, 4), dtype=np.float32)
        else:
            p.accumulate = 1
        ids = np.zeros((height, width), dtype=np.uint32) if want_ids else None
        rgba = np.zeros((height, width, 4), dtype=np.uint8) if want_rgba else None
        cnt = Counters()
        rc = self.l.orc_render(self.h, C.byref(p), _ptr(fd), _ptr(accum), _ptr(ids) if want_ids else None,
                               _ptr(rgba) if want_rgba else None, C.byref(cnt))
        assert rc == 0, "orc_render failed: %d" % rc
        return accum, ids, rgba, cnt


def resolve(frame_data, accum, integrator=PATH, seed=0, frame_index=0, fast=False):
    h, w = accum.shape[:2]
    fd = np.frombuffer(bytes(frame_data), dtype=np.uint8).copy()
    p = Params(width=w, height=h, integrator=integrator, seed=seed, frame_index=frame_index)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    accum = np.ascontiguousarray(accum, dtype=np.float32)
    assert lib(fast).orc_resolve(C.byref(p), _ptr(fd), _ptr(accum), _ptr(out)) == 0
    return out


def present(binding0, binding1, out_w, out_h, fast=False):
    """Fullscreen.frag on two rgba8 images (h, w, 4) -> (out_h, out_w, 4) framebuffer, row 0 = top."""
    b0 = np.ascontiguousarray(binding0, dtype=np.uint8)
    b1 = np.ascontiguousarray(binding1, dtype=np.uint8)
    th, tw = b0.shape[:2]
    out = np.zeros((out_h, out_w, 4), dtype=np.uint8)
    assert lib(fast).orc_present(_ptr(b0), _ptr(b1), tw, th, _ptr(out), out_w, out_h) == 0
    return out


# ---- the reference's own Camera.cpp, compiled from /root/reference into oracle/_ref ------------
def ref_camera_lib():
    """Returns the ctypes handle of oracle/_ref/libref_camera.so or None when it was never built."""
    path = os.path.join(_HERE, "_ref", "libref_camera.so")
    if not os.path.exists(path):
        return None
    l = C.CDLL(path)
    l.ref_camera_update.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_void_p]
    l.ref_default_frame_data.argtypes = [C.c_void_p]
    l.ref_camera_move.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float]
    return l
