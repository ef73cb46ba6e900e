"""ctypes binding of libvkrt_cuda.so -- exactly the entry points include/vkrt.h declares.

There is no CPU fallback: if the shared library is missing it is built with nvcc, and if that
fails the import raises.  Nothing here touches oracle/.
"""
import ctypes as C
import os

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvkrt_cuda.so")

SUCCESS, NO_SUITABLE_GPU, NO_SUITABLE_SURFACE, UNKNOWN, CUDA_ERROR, NCCL_ERROR, BAD_ARG = range(7)
ERROR_NAMES = ["SUCCESS", "NO_SUITABLE_GPU", "NO_SUITABLE_SURFACE", "UNKNOWN", "CUDA_ERROR", "NCCL_ERROR", "BAD_ARG"]

INTEGRATOR_WHITTED, INTEGRATOR_PATH = 0, 1
VARIANT_MEGAKERNEL, VARIANT_WAVEFRONT = 0, 1
SCENE_TRACER, SCENE_RAYTRACER = 0, 1
FLAG_PROGRESSIVE, FLAG_HIT_IDS, FLAG_STATS, FLAG_NO_RESOLVE, FLAG_SERIAL_WAVES, FLAG_LAUNCH_TIMING = 1, 2, 4, 8, 16, 32
MAT_DIFFUSE, MAT_DIELECTRIC = 0, 1
TILING_LINEAR, TILING_OPTIMAL = 0, 1
SEMAPHORE_ACQUIRE, SEMAPHORE_RELEASE = 0, 1


class Vec3a(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("_pad", C.c_float)]


class CameraData(C.Structure):  # ref: Include/Camera.h:5-12
    _fields_ = [("pos", Vec3a), ("dir", Vec3a), ("right", Vec3a), ("up", Vec3a)]


class FrameData(C.Structure):  # ref: Include/GraphicsDevice.h:20-29
    _fields_ = [("aspect_ratio", C.c_float), ("seed", C.c_float), ("_pad0", C.c_float * 2),
                ("light_pos", Vec3a), ("camera", CameraData)]


class Triangle(C.Structure):  # ref: Include/GraphicsDevice.h:13-18
    _fields_ = [("v0", Vec3a), ("v1", Vec3a), ("v2", Vec3a)]


class Material(C.Structure):
    _fields_ = [("albedo", C.c_float * 3), ("roughness", C.c_float), ("emissive", C.c_float * 3),
                ("metalness", C.c_float), ("type", C.c_uint32), ("_pad", C.c_uint32 * 3)]


class CreateInfo(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("spp", C.c_uint32), ("max_depth", C.c_uint32), ("integrator", C.c_uint32),
                ("variant", C.c_uint32), ("frames_in_flight", C.c_uint32), ("device_id", C.c_int32),
                ("flags", C.c_uint32), ("tile_shard_rank", C.c_uint32), ("tile_shard_count", C.c_uint32),
                ("sample_shard_rank", C.c_uint32), ("sample_shard_count", C.c_uint32), ("stream", C.c_void_p),
                ("n_devices", C.c_uint32), ("device_ids", C.c_int32 * 8)]


class ExchangeHandle(C.Structure):  # include/vkrt.h: the frame-exchange block of the gathering context (CUDA IPC)
    _fields_ = [("ipc", C.c_uint8 * 64), ("bytes", C.c_uint64), ("width", C.c_uint32), ("height", C.c_uint32),
                ("tile_shard_count", C.c_uint32), ("sample_shard_count", C.c_uint32), ("magic", C.c_uint32), ("_pad", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("node_visits", C.c_uint64),
                ("leaf_tests", C.c_uint64), ("paths", C.c_uint64), ("frames", C.c_uint64),
                ("shared_primary_rays", C.c_uint64), ("zero_term_shadow_rays", C.c_uint64)]


class BvhInfo(C.Structure):
    _fields_ = [("n_spheres", C.c_uint32), ("n_nodes", C.c_uint32), ("node_bytes", C.c_uint32),
                ("build_ms", C.c_float), ("build_launches", C.c_uint32), ("depth", C.c_uint32),
                ("n_triangles", C.c_uint32), ("n_tri_nodes", C.c_uint32), ("tri_depth", C.c_uint32), ("tri_build_ms", C.c_float),
                ("traversal_depth", C.c_uint32), ("traversal_is_sah", C.c_uint32), ("traversal_build_ms", C.c_float)]


class ExternalImage(C.Structure):  # include/vkrt.h: one exported traced image (ref: Source/GraphicsDevice.cpp:664-699)
    _fields_ = [("struct_size", C.c_uint32), ("fd", C.c_int32), ("allocation_size", C.c_uint64), ("offset", C.c_uint64),
                ("tiling", C.c_uint32), ("row_pitch", C.c_uint32), ("dedicated", C.c_uint32), ("_pad", C.c_uint32)]


assert C.sizeof(ExternalImage) == 40 and C.sizeof(ExchangeHandle) == 96
assert C.sizeof(CameraData) == 64 and C.sizeof(FrameData) == 96 and C.sizeof(Triangle) == 48
assert C.sizeof(Material) == 48
assert FrameData.seed.offset == 4 and FrameData.light_pos.offset == 16 and FrameData.camera.offset == 32

_vp, _u32, _i32, _sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
_P = C.POINTER

# name -> (argtypes, restype); every symbol include/vkrt.h declares
SIGNATURES = {
    "vkrt_create": ([_P(CreateInfo), _P(_vp)], C.c_int8),
    "vkrt_destroy": ([_vp], C.c_int8),
    "vkrt_draw": ([_vp, _P(FrameData)], C.c_int8),
    "vkrt_wait_idle": ([_vp], C.c_int8),
    "vkrt_flush": ([_vp], C.c_int8),
    "vkrt_set_sampling": ([_vp, _u32, _u32], C.c_int8),
    "vkrt_set_seed": ([_vp, C.c_uint64], C.c_int8),
    "vkrt_set_frame_index": ([_vp, _u32], C.c_int8),
    "vkrt_reset_accum": ([_vp], C.c_int8),
    "vkrt_set_triangles": ([_vp, _vp, _u32], C.c_int8),
    "vkrt_set_triangle_material": ([_vp, _u32], C.c_int8),
    "vkrt_set_triangle_materials": ([_vp, _vp, _u32], C.c_int8),
    "vkrt_load_obj": ([C.c_char_p, _P(C.c_float), _vp, _u32, _P(_u32)], C.c_int8),
    "vkrt_set_materials": ([_vp, _vp, _u32], C.c_int8),
    "vkrt_set_spheres": ([_vp, _vp, _vp, _u32], C.c_int8),
    "vkrt_set_planes": ([_vp, _vp, _vp, _u32], C.c_int8),
    "vkrt_use_default_scene": ([_vp, _u32], C.c_int8),
    "vkrt_build_bvh": ([_vp], C.c_int8),
    "vkrt_clear_bvh": ([_vp], C.c_int8),
    "vkrt_get_rgba8": ([_vp, _P(_vp), _P(_sz)], C.c_int8),
    "vkrt_get_accum": ([_vp, _P(_vp)], C.c_int8),
    "vkrt_get_hit_ids": ([_vp, _P(_vp)], C.c_int8),
    "vkrt_get_stream": ([_vp, _P(_vp)], C.c_int8),
    "vkrt_read_rgba8": ([_vp, _vp, _sz], C.c_int8),
    "vkrt_read_accum": ([_vp, _vp, _sz], C.c_int8),
    "vkrt_read_hit_ids": ([_vp, _vp, _sz], C.c_int8),
    "vkrt_read_rgba8_async": ([_vp, _vp, _sz], C.c_int8),
    "vkrt_resolve": ([_vp], C.c_int8),
    "vkrt_present": ([_vp, _vp, _u32, _u32], C.c_int8),
    "vkrt_import_vk_image": ([_vp, _u32, _P(ExternalImage)], C.c_int8),
    "vkrt_bind_rgba8_target": ([_vp, _u32, _vp, _sz], C.c_int8),
    "vkrt_debug_bind_array_target": ([_vp, _u32], C.c_int8),
    "vkrt_import_vk_semaphore": ([_vp, _u32, _u32, C.c_int32, _u32], C.c_int8),
    "vkrt_release_external": ([_vp], C.c_int8),
    "vkrt_get_counters": ([_vp, _P(Counters)], C.c_int8),
    "vkrt_reset_counters": ([_vp], C.c_int8),
    "vkrt_last_frame_timing": ([_vp, _P(C.c_float), _P(C.c_float), _P(_u32)], C.c_int8),
    "vkrt_last_frame_traversal_timing": ([_vp, _P(C.c_float), _P(_u32), _P(C.c_float)], C.c_int8),
    "vkrt_debug_dump_timeline": ([_vp, C.c_char_p], C.c_int8),
    "vkrt_get_bvh_info": ([_vp, _P(BvhInfo)], C.c_int8),
    "vkrt_read_bvh_nodes": ([_vp, _vp, _sz], C.c_int8),
    "vkrt_read_bvh_traversal_nodes": ([_vp, _vp, _sz], C.c_int8),
    "vkrt_read_bvh_qnodes": ([_vp, _vp, _sz, _P(C.c_float)], C.c_int8),
    "vkrt_pack_shard": ([_vp, _P(_vp), _P(_sz)], C.c_int8),
    "vkrt_pack_shard_into": ([_vp, _vp, _sz], C.c_int8),
    "vkrt_shard_floats": ([_vp, _u32, _P(_sz)], C.c_int8),
    "vkrt_unpack_shard": ([_vp, _vp, _u32, _u32, _i32], C.c_int8),
    "vkrt_exchange_create": ([_vp, _P(ExchangeHandle)], C.c_int8),
    "vkrt_exchange_open": ([_vp, _P(ExchangeHandle)], C.c_int8),
    "vkrt_exchange_attach": ([_vp, _vp], C.c_int8),
    "vkrt_exchange_close": ([_vp], C.c_int8),
    "vkrt_measure_fp32_peak": ([_i32, _P(C.c_float)], C.c_int8),
    "vkrt_measure_l2_bandwidth": ([_i32, _P(C.c_float)], C.c_int8),
    "vkrt_last_error_string": ([_vp], C.c_char_p),
    "vkrt_version": ([], C.c_char_p),
}

_LIB = None


def load():
    """Loads (building first if needed) libvkrt_cuda.so.  Raises when it cannot: no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("VKRT_LIB") or LIB_PATH      # VKRT_LIB: a tuning build of the same library
    if not os.path.exists(path):
        if path != LIB_PATH:
            raise OSError("VKRT_LIB=%s does not exist" % path)
        _build.build()
    lib = C.CDLL(path)
    for name, (args, res) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.argtypes = args
        fn.restype = res
    _LIB = lib
    return lib


class VkrtError(RuntimeError):
    def __init__(self, code, msg):
        self.code = code
        name = ERROR_NAMES[code] if 0 <= code < len(ERROR_NAMES) else str(code)
        super().__init__("%s: %s" % (name, msg))


def check(lib, ctx, rc):
    if rc != SUCCESS:
        msg = lib.vkrt_last_error_string(ctx)
        raise VkrtError(rc, msg.decode() if msg else "")
