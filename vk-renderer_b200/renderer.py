"""Thin object wrapper over the C ABI (one vkrt_ctx).  numpy in, numpy out; device pointers are
exposed for the torch.distributed plumbing in sharding.py."""
import ctypes as C

import numpy as np

from . import _lib as L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_materials(albedo, roughness, emissive, metalness, mtype):
    """-> (n, 12) float32 array in the vkrt_material layout (type stored as uint32 bits)."""
    n = len(roughness)
    m = np.zeros((n, 12), dtype=np.float32)
    m[:, 0:3] = albedo
    m[:, 3] = roughness
    m[:, 4:7] = emissive
    m[:, 7] = metalness
    m.view(np.uint32)[:, 8] = np.asarray(mtype, dtype=np.uint32)
    return m


class Renderer:
    def __init__(self, width, height, spp=4, max_depth=0, integrator=L.INTEGRATOR_PATH, variant=L.VARIANT_MEGAKERNEL,
                 flags=0, device_id=0, frames_in_flight=2, tile_shard=(0, 1), sample_shard=(0, 1), stream=None, device_ids=None):
        """device_ids: in-library multi-GPU -- ONE context renders every frame on these CUDA devices (interleaved tile
        shards, frame exchange over peer memory into device_ids[0]); tile_shard / sample_shard stay at their defaults."""
        self.lib = L.load()
        info = L.CreateInfo(struct_size=C.sizeof(L.CreateInfo), width=width, height=height, spp=spp, max_depth=max_depth,
                            integrator=integrator, variant=variant, frames_in_flight=frames_in_flight,
                            device_id=device_id, flags=flags, tile_shard_rank=tile_shard[0],
                            tile_shard_count=tile_shard[1], sample_shard_rank=sample_shard[0],
                            sample_shard_count=sample_shard[1], stream=stream)
        if device_ids is not None:
            info.n_devices = len(device_ids)
            for i, d in enumerate(device_ids):
                info.device_ids[i] = d
            device_id = device_ids[0]
        ctx = C.c_void_p()
        rc = self.lib.vkrt_create(C.byref(info), C.byref(ctx))
        L.check(self.lib, None, rc)
        self.ctx = ctx
        self.width, self.height, self.device_id = width, height, device_id
        self.integrator = integrator

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.vkrt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        L.check(self.lib, self.ctx, rc)

    # ---- scene --------------------------------------------------------------------------------
    def use_default_scene(self, which):
        self._ck(self.lib.vkrt_use_default_scene(self.ctx, which))
        return self

    def set_materials(self, mats):
        mats = np.ascontiguousarray(mats)
        assert mats.ndim == 2 and mats.shape[1] * mats.dtype.itemsize == 48
        self._ck(self.lib.vkrt_set_materials(self.ctx, _ptr(mats), mats.shape[0]))

    def set_spheres(self, spheres, mat_id):
        spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
        mat_id = np.ascontiguousarray(mat_id, dtype=np.uint32)
        assert mat_id.shape[0] == spheres.shape[0]
        self._ck(self.lib.vkrt_set_spheres(self.ctx, _ptr(spheres), _ptr(mat_id), spheres.shape[0]))

    def set_planes(self, planes, mat_id):
        planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 4)
        mat_id = np.ascontiguousarray(mat_id, dtype=np.uint32)
        self._ck(self.lib.vkrt_set_planes(self.ctx, _ptr(planes), _ptr(mat_id), planes.shape[0]))

    def set_triangles(self, tris, mat_id=None):
        tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 12)
        self._ck(self.lib.vkrt_set_triangles(self.ctx, _ptr(tris), tris.shape[0]))
        if mat_id is not None:
            self._ck(self.lib.vkrt_set_triangle_material(self.ctx, mat_id))

    def set_triangle_materials(self, mat_ids):
        """One material id per triangle (None / empty: every triangle uses the shared material again)."""
        m = np.ascontiguousarray(mat_ids if mat_ids is not None else [], dtype=np.uint32)
        self._ck(self.lib.vkrt_set_triangle_materials(self.ctx, _ptr(m), m.shape[0]))

    def set_scene(self, scene):
        """scene: scenes.Scene (materials / spheres / planes / triangles as numpy arrays)."""
        self.set_materials(scene.materials)
        self.set_spheres(scene.spheres, scene.sphere_mat)
        self.set_planes(scene.planes, scene.plane_mat)
        self.set_triangles(scene.triangles, scene.tri_mat)
        if getattr(scene, "tri_mats", None) is not None:
            self.set_triangle_materials(scene.tri_mats)
        return self

    def build_bvh(self):
        self._ck(self.lib.vkrt_build_bvh(self.ctx))
        return self.bvh_info()

    def clear_bvh(self):
        self._ck(self.lib.vkrt_clear_bvh(self.ctx))

    def bvh_info(self):
        info = L.BvhInfo()
        self._ck(self.lib.vkrt_get_bvh_info(self.ctx, C.byref(info)))
        return info

    def bvh_nodes(self):
        n = self.bvh_info().n_nodes
        out = np.zeros((n, 16), dtype=np.float32)
        if n:
            self._ck(self.lib.vkrt_read_bvh_nodes(self.ctx, _ptr(out), out.nbytes))
        return out

    def bvh_traversal_nodes(self):
        """The exact nodes of the tree the 32-byte traversal nodes are made from (binned SAH; same layout as bvh_nodes)."""
        n = self.bvh_info().n_nodes
        out = np.zeros((n, 16), dtype=np.float32)
        if n:
            self._ck(self.lib.vkrt_read_bvh_traversal_nodes(self.ctx, _ptr(out), out.nbytes))
        return out

    def bvh_qnodes(self):
        """-> ((n, 8) uint32 traversal nodes, grid[6] float32): two child records {x: lo | hi << 16, y, z, ref}."""
        n = self.bvh_info().n_nodes
        out = np.zeros((n, 8), dtype=np.uint32)
        grid = (C.c_float * 6)()
        if n:
            self._ck(self.lib.vkrt_read_bvh_qnodes(self.ctx, _ptr(out), out.nbytes, grid))
        return out, np.array(list(grid), dtype=np.float32)

    # ---- frame --------------------------------------------------------------------------------
    def set_sampling(self, spp, max_depth):
        self._ck(self.lib.vkrt_set_sampling(self.ctx, spp, max_depth))

    def set_seed(self, seed):
        self._ck(self.lib.vkrt_set_seed(self.ctx, seed))

    def set_frame_index(self, i):
        self._ck(self.lib.vkrt_set_frame_index(self.ctx, i))

    def reset_accum(self):
        self._ck(self.lib.vkrt_reset_accum(self.ctx))

    def draw(self, frame_data):
        if not isinstance(frame_data, L.FrameData):
            frame_data = L.FrameData.from_buffer_copy(bytes(frame_data))
        self._ck(self.lib.vkrt_draw(self.ctx, C.byref(frame_data)))

    def flush(self):
        self._ck(self.lib.vkrt_flush(self.ctx))

    def wait_idle(self):
        self._ck(self.lib.vkrt_wait_idle(self.ctx))

    def resolve(self):
        self._ck(self.lib.vkrt_resolve(self.ctx))

    # ---- outputs ------------------------------------------------------------------------------
    def read_rgba8(self):
        out = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        self._ck(self.lib.vkrt_read_rgba8(self.ctx, _ptr(out), out.nbytes))
        return out

    def read_rgba8_async(self, pinned_ptr, nbytes):
        self._ck(self.lib.vkrt_read_rgba8_async(self.ctx, pinned_ptr, nbytes))

    def present(self, out_w, out_h):
        """Fullscreen.frag on the two traced image slots -> (out_h, out_w, 4) uint8 framebuffer, row 0 = top."""
        out = np.zeros((out_h, out_w, 4), dtype=np.uint8)
        self._ck(self.lib.vkrt_present(self.ctx, _ptr(out), out_w, out_h))
        return out

    def read_accum(self):
        out = np.zeros((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.vkrt_read_accum(self.ctx, _ptr(out), out.nbytes))
        return out

    def read_hit_ids(self):
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        self._ck(self.lib.vkrt_read_hit_ids(self.ctx, _ptr(out), out.nbytes))
        return out

    def counters(self):
        c = L.Counters()
        self._ck(self.lib.vkrt_get_counters(self.ctx, C.byref(c)))
        return c

    def reset_counters(self):
        self._ck(self.lib.vkrt_reset_counters(self.ctx))

    def last_frame_timing(self):
        a, b, n = C.c_float(), C.c_float(), C.c_uint32()
        self._ck(self.lib.vkrt_last_frame_timing(self.ctx, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def last_frame_traversal_timing(self):
        a, n, t = C.c_float(), C.c_uint32(), C.c_float()
        self._ck(self.lib.vkrt_last_frame_traversal_timing(self.ctx, C.byref(a), C.byref(n), C.byref(t)))
        return a.value, n.value, t.value

    def dump_timeline(self, path):
        self._ck(self.lib.vkrt_debug_dump_timeline(self.ctx, path.encode()))

    def stream_ptr(self):
        s = C.c_void_p()
        self._ck(self.lib.vkrt_get_stream(self.ctx, C.byref(s)))
        return s.value or 0

    def accum_ptr(self):
        p = C.c_void_p()
        self._ck(self.lib.vkrt_get_accum(self.ctx, C.byref(p)))
        return p.value

    def rgba8_ptr(self):
        p, pitch = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.vkrt_get_rgba8(self.ctx, C.byref(p), C.byref(pitch)))
        return p.value, pitch.value

    # ---- Vulkan <-> CUDA interop (the traced images the engine presents) ----------------------------
    def import_vk_image(self, slot, fd, allocation_size, offset=0, tiling=L.TILING_OPTIMAL, row_pitch=0, dedicated=False):
        """traced_images[slot] exported by the engine with vkGetMemoryFdKHR (ref: Source/GraphicsDevice.cpp:664-699)."""
        im = L.ExternalImage(struct_size=C.sizeof(L.ExternalImage), fd=fd, allocation_size=allocation_size, offset=offset,
                             tiling=tiling, row_pitch=row_pitch, dedicated=1 if dedicated else 0)
        self._ck(self.lib.vkrt_import_vk_image(self.ctx, slot, C.byref(im)))

    def bind_rgba8_target(self, slot, dev_ptr, row_pitch=0):
        self._ck(self.lib.vkrt_bind_rgba8_target(self.ctx, slot, C.c_void_p(dev_ptr), row_pitch))

    def debug_bind_array_target(self, slot):
        self._ck(self.lib.vkrt_debug_bind_array_target(self.ctx, slot))

    def import_vk_semaphore(self, slot, which, fd, timeline=False):
        self._ck(self.lib.vkrt_import_vk_semaphore(self.ctx, slot, which, fd, 1 if timeline else 0))

    def release_external(self):
        self._ck(self.lib.vkrt_release_external(self.ctx))

    # ---- sharding -----------------------------------------------------------------------------
    def pack_shard(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.vkrt_pack_shard(self.ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def pack_shard_into(self, dev_ptr, n_floats):
        self._ck(self.lib.vkrt_pack_shard_into(self.ctx, C.c_void_p(dev_ptr), n_floats))

    def shard_floats(self, tile_rank=0):
        n = C.c_size_t()
        self._ck(self.lib.vkrt_shard_floats(self.ctx, tile_rank, C.byref(n)))
        return n.value

    def unpack_shard(self, dev_ptr, tile_rank, tile_count, add=False):
        self._ck(self.lib.vkrt_unpack_shard(self.ctx, C.c_void_p(dev_ptr), tile_rank, tile_count, 1 if add else 0))

    # ---- frame exchange over peer memory ------------------------------------------------------------
    def exchange_create(self):
        """Gathering context (tile rank 0, sample rank 0): -> the handle (bytes) the other processes open."""
        h = L.ExchangeHandle()
        self._ck(self.lib.vkrt_exchange_create(self.ctx, C.byref(h)))
        return bytes(h)

    def exchange_open(self, handle_bytes):
        h = L.ExchangeHandle.from_buffer_copy(handle_bytes)
        self._ck(self.lib.vkrt_exchange_open(self.ctx, C.byref(h)))

    def exchange_attach(self, gathering_renderer):
        self._ck(self.lib.vkrt_exchange_attach(self.ctx, gathering_renderer.ctx))

    def exchange_close(self):
        self._ck(self.lib.vkrt_exchange_close(self.ctx))


def measure_fp32_peak(device_id=0):
    lib = L.load()
    v = C.c_float()
    L.check(lib, None, lib.vkrt_measure_fp32_peak(device_id, C.byref(v)))
    return v.value


def measure_l2_bandwidth(device_id=0):
    lib = L.load()
    v = C.c_float()
    L.check(lib, None, lib.vkrt_measure_l2_bandwidth(device_id, C.byref(v)))
    return v.value


def load_obj(path, xform=None):
    """vkrt_load_obj: the triangles of a Wavefront OBJ file as an (n, 12) float32 array in the reference's 48-byte
    Triangle layout (xform: optional 3x4 row-major matrix applied to every vertex)."""
    lib = L.load()
    n = C.c_uint32()
    m = None
    if xform is not None:
        m = (C.c_float * 12)(*[float(x) for x in np.asarray(xform, dtype=np.float32).reshape(12)])
    L.check(lib, None, lib.vkrt_load_obj(path.encode(), m, None, 0, C.byref(n)))
    out = np.zeros((n.value, 12), dtype=np.float32)
    if n.value:
        L.check(lib, None, lib.vkrt_load_obj(path.encode(), m, _ptr(out), n.value, C.byref(n)))
    return out
