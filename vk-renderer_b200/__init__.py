"""vk-renderer_b200: B200-native ray-tracing hot path of stickyfingies/vk-renderer.

libvkrt_cuda.so (csrc/, hand-written CUDA for sm_100a behind the C ABI of include/vkrt.h) plus
the host-side mirror of the reference's engine interface (device.py).  No CPU fallback: importing
`_lib` builds/loads the CUDA library or raises.
"""
from . import _lib
from ._lib import (FLAG_HIT_IDS, FLAG_LAUNCH_TIMING, FLAG_NO_RESOLVE, FLAG_PROGRESSIVE, FLAG_SERIAL_WAVES, FLAG_STATS, INTEGRATOR_PATH, INTEGRATOR_WHITTED,
                   SCENE_RAYTRACER, SCENE_TRACER, SEMAPHORE_ACQUIRE, SEMAPHORE_RELEASE, TILING_LINEAR, TILING_OPTIMAL,
                   VARIANT_MEGAKERNEL, VARIANT_WAVEFRONT, FrameData, VkrtError)
from .device import Camera, GraphicsDevice, default_camera, default_frame_data
from .renderer import Renderer, load_obj, measure_fp32_peak, measure_l2_bandwidth, pack_materials
from . import scenes

__all__ = ["Renderer", "GraphicsDevice", "Camera", "FrameData", "default_camera", "default_frame_data", "scenes",
           "VkrtError", "pack_materials", "load_obj", "measure_fp32_peak", "measure_l2_bandwidth"]
