"""Host-side mirror of the reference's engine interface for the hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference:
  * GraphicsDevice.Construct / Destruct / Draw / WaitIdle  -- Include/GraphicsDevice.h:41-97,
    Source/GraphicsDevice.cpp:222 (Construct), :1138 (Destruct), :1215 (Draw), :1344 (WaitIdle)
  * GraphicsDevice.Error                                    -- Include/GraphicsDevice.h:46-52
  * Camera / CameraData / FrameData                         -- Include/Camera.h, Source/Camera.cpp
The window / swapchain / present half of the reference is out of scope: the
"window" here is just an extent used for aspect_ratio, as the swapchain extent is at
Source/GraphicsDevice.cpp:1260.
"""
import ctypes as C
import ctypes.util
import enum
import math
import time

import numpy as np

from . import _lib as L
from .renderer import Renderer

_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.cosf.argtypes = _libm.sinf.argtypes = [C.c_float]
_libm.cosf.restype = _libm.sinf.restype = C.c_float
_libc = C.CDLL(ctypes.util.find_library("c") or "libc.so.6")
_libc.rand.restype = C.c_int
_RAND_MAX = 2147483647

_f = np.float32


def _v3(x, y, z):
    return np.array([x, y, z], dtype=np.float32)


def _normalize(v):
    # glm::normalize = v * inversesqrt(dot(v,v)); glm::dot sums (x*x + y*y) + z*z in binary32
    d = _f(_f(_f(v[0] * v[0]) + _f(v[1] * v[1])) + _f(v[2] * v[2]))
    inv = _f(_f(1.0) / np.sqrt(d, dtype=np.float32))
    return (v * inv).astype(np.float32)


def _cross(x, y):
    # glm::cross(x, y)
    return _v3(_f(x[1] * y[2]) - _f(y[1] * x[2]), _f(x[2] * y[0]) - _f(y[2] * x[0]), _f(x[0] * y[1]) - _f(y[0] * x[1]))


class Camera:
    """Source/Camera.cpp: yaw/pitch -> (pos, dir, right, up), all binary32 like glm."""

    def __init__(self):
        self.pos = _v3(0.0, 64.0, -48.0)           # Camera.cpp:6
        self.front = _v3(0.0, 0.0, -1.0)
        self.pitch, self.yaw = _f(0.0), _f(90.0)   # Camera.cpp:7
        self.dir = self.right = self.up = None
        self.update()

    def update(self):                               # Camera.cpp:48-64
        self.pitch = _f(min(max(float(self.pitch), -89.0), 89.0))
        rad = _f(0.01745329251994329576923690768489)          # glm::radians
        p, y = _f(_f(self.pitch) * rad), _f(_f(self.yaw) * rad)
        # unqualified cos()/sin() on a float resolve to the C double functions in Camera.cpp:54-56, so the
        # products are formed in binary64 and rounded once on assignment to the vec3
        cp, sp = math.cos(float(p)), math.sin(float(p))
        cy, sy = math.cos(float(y)), math.sin(float(y))
        front = _v3(_f(cp * cy), _f(sp), _f(cp * sy))
        self.front = _normalize(front)
        self.dir = _normalize((self.pos + self.front).astype(np.float32) - self.pos)
        up = _v3(0.0, 1.0, 0.0)
        self.right = _normalize(_cross(up, self.dir))
        self.up = _cross(self.dir, self.right)

    def _move(self, v, speed):
        self.pos = (self.pos + v * _f(speed)).astype(np.float32)
        self.update()

    def move_forward(self, speed): self._move(self.front, speed)      # Camera.cpp:12-16
    def move_backward(self, speed): self._move(-self.front, speed)
    def move_left(self, speed): self._move(-self.right, speed)
    def move_right(self, speed): self._move(self.right, speed)
    def move_up(self, speed): self._move(self.up, speed)
    def move_down(self, speed): self._move(-self.up, speed)

    @property
    def data(self):
        cd = L.CameraData()
        for name in ("pos", "dir", "right", "up"):
            v = getattr(self, name)
            setattr(cd, name, L.Vec3a(float(v[0]), float(v[1]), float(v[2]), 0.0))
        return cd


def default_camera():
    """The view Source/Main.cpp:134-139 sets before the first Draw."""
    cam = Camera()
    cam.pos = _v3(32.8509, 30.6991, -106.389)
    cam.pitch, cam.yaw = _f(4.44998), _f(-602.79)
    cam.update()
    return cam


def default_frame_data(aspect_ratio=1024.0 / 768.0, seed=0.0, camera=None):
    """FrameData as Main.cpp:141-142 fills it (light_pos = (0,64,0)), with aspect/seed as Draw would."""
    cam = camera or default_camera()
    fd = L.FrameData()
    fd.aspect_ratio, fd.seed = aspect_ratio, seed
    fd.light_pos = L.Vec3a(0.0, 64.0, 0.0, 0.0)
    fd.camera = cam.data
    return fd


class GraphicsDevice:
    """Drop-in for the reference's GraphicsDevice on the ray-tracing path (no window, no present)."""

    class Error(enum.IntEnum):            # Include/GraphicsDevice.h:46-52 (+ the C ABI's extra codes)
        SUCCESS = 0
        NO_SUITABLE_GPU = 1
        NO_SUITABLE_SURFACE = 2
        UNKNOWN = 3
        CUDA_ERROR = 4
        NCCL_ERROR = 5
        BAD_ARG = 6

    class CreateInfo:                      # Include/GraphicsDevice.h:57-72
        def __init__(self, window=None, swapchainSize=3, framesInFlight=2, raytrace_resolution=1024, debug=False,
                     # generalisations of the shader's compile-time constants / single square target:
                     width=None, height=None, spp=4, max_depth=None, integrator=L.INTEGRATOR_PATH,
                     variant=L.VARIANT_MEGAKERNEL, scene=L.SCENE_TRACER, swapchain_extent=(1024, 768), flags=0,
                     device_id=0):
            self.window, self.swapchainSize, self.framesInFlight = window, swapchainSize, framesInFlight
            self.raytrace_resolution, self.debug = raytrace_resolution, debug
            self.width = width or raytrace_resolution
            self.height = height or raytrace_resolution
            self.spp, self.max_depth, self.integrator, self.variant = spp, max_depth or 0, integrator, variant
            self.scene, self.swapchain_extent, self.flags, self.device_id = scene, swapchain_extent, flags, device_id

    def __init__(self):
        self.renderer = None
        self._extent = (1024, 768)

    def Construct(self, info):
        try:
            self.renderer = Renderer(info.width, info.height, spp=info.spp, max_depth=info.max_depth,
                                     integrator=info.integrator, variant=info.variant, flags=info.flags,
                                     device_id=info.device_id, frames_in_flight=info.framesInFlight)
            if info.scene is not None:
                self.renderer.use_default_scene(info.scene)
        except L.VkrtError as e:
            print("[app] - err :: Graphics device creation failed :: %s" % e)     # Main.cpp:121
            return GraphicsDevice.Error(e.code)
        self._extent = info.swapchain_extent
        _libc.srand(int(time.time()))                                              # GraphicsDevice.cpp:1133
        return GraphicsDevice.Error.SUCCESS

    def Destruct(self):
        if self.renderer is None:
            return GraphicsDevice.Error.UNKNOWN
        self.renderer.close()
        self.renderer = None
        return GraphicsDevice.Error.SUCCESS

    def Draw(self, frame_data):
        """void in the reference; errors are swallowed there too (VkResults ignored in Draw)."""
        fd = L.FrameData.from_buffer_copy(bytes(frame_data))                       # frame_data_real = frame_data (:1258)
        fd.aspect_ratio = float(self._extent[0]) / float(self._extent[1])          # :1260
        fd.seed = float(_libc.rand()) / float(_RAND_MAX)                           # :1262
        self.renderer.draw(fd)

    def WaitIdle(self):
        self.renderer.wait_idle()
