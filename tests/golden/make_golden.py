#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ (run in the build container, where
/root/reference exists):

  camera_poses.json   -- CameraData produced by the REFERENCE'S OWN Source/Camera.cpp (compiled from
                         /root/reference into oracle/_ref by oracle/Makefile) for the default view of
                         Source/Main.cpp:134-139 and a set of seeded poses / moves
  frames.npz          -- small frames rendered by the CPU oracle (oracle/vkrt_oracle.cpp): accumulator,
                         primary hit ids, rgba8 and ray counts for both integrators and for an LBVH scene

The reference ships no golden vectors of its own and its shaders cannot run here, so frames.npz pins the
oracle against regressions (and the GPU against the oracle at rest); camera_poses.json is a true
reference-generated vector.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
import vk_renderer_b200 as V  # noqa: E402
from helpers import apply_scene  # noqa: E402

FRAMES = {
    # name: (scene, w, h, spp, depth, integrator, sphere_mode, seed, frame_seed)
    "whitted_default": ("raytracer", 64, 48, 1, 2, O.WHITTED, O.LITERAL, 0, 0.0),
    "path_default": ("tracer", 64, 48, 4, 4, O.PATH, O.LITERAL, 7, 0.25),
    "path_default_d8": ("tracer", 48, 27, 16, 8, O.PATH, O.LITERAL, 2026, 0.5),
    "path_random64_bvh": ("random64", 48, 36, 2, 8, O.PATH, O.S_BVH, 7, 0.75),
}


def scene_for(name):
    if name == "raytracer":
        return V.scenes.raytracer_default()
    if name == "tracer":
        return V.scenes.tracer_default()
    return V.scenes.random_spheres(64)


def main():
    O.build()
    rl = O.ref_camera_lib()
    assert rl is not None, "oracle/_ref/libref_camera.so missing: run `make -C oracle` where /root/reference exists"
    poses = []
    fd = np.zeros(24, dtype=np.float32)
    rl.ref_default_frame_data(fd.ctypes.data_as(C.c_void_p))
    default = {"light_pos": fd[4:7].tolist(), "camera": fd[8:24].reshape(4, 4)[:, :3].tolist()}
    rng = np.random.default_rng(20261017)
    for _ in range(64):
        pos = rng.uniform(-100, 100, 3).astype(np.float32)
        pitch, yaw = np.float32(rng.uniform(-120, 120)), np.float32(rng.uniform(-720, 720))
        out = np.zeros(16, dtype=np.float32)
        rl.ref_camera_update(pos.ctypes.data_as(C.POINTER(C.c_float)), float(pitch), float(yaw), out.ctypes.data_as(C.c_void_p))
        op, speed = int(rng.integers(0, 6)), float(np.float32(rng.uniform(0.1, 5.0)))
        moved = out.copy()
        rl.ref_camera_move(moved.ctypes.data_as(C.c_void_p), float(pitch), float(yaw), op, speed)
        poses.append({"pos": pos.tolist(), "pitch": float(pitch), "yaw": float(yaw),
                      "camera": out.reshape(4, 4)[:, :3].tolist(), "move_op": op, "move_speed": speed,
                      "camera_after_move": moved.reshape(4, 4)[:, :3].tolist()})
    with open(os.path.join(HERE, "camera_poses.json"), "w") as f:
        json.dump({"generator": "reference Source/Camera.cpp via oracle/_ref/libref_camera.so",
                   "default_view": default, "poses": poses}, f)

    out = {}
    for name, (scn, w, h, spp, depth, integ, mode, seed, fseed) in FRAMES.items():
        scene = scene_for(scn)
        sc = apply_scene(O, scene)
        if mode == O.S_BVH:
            sc.build_bvh()
        fdata = V.default_frame_data(aspect_ratio=w / h, seed=fseed)
        acc, ids, rgba, cnt = sc.render(fdata, w, h, spp=spp, max_depth=depth, integrator=integ, sphere_mode=mode, seed=seed)
        out[name + ".accum"] = acc
        out[name + ".ids"] = ids
        out[name + ".rgba"] = rgba
        out[name + ".counts"] = np.array([cnt.closest_rays, cnt.shadow_rays, cnt.paths], dtype=np.uint64)
        out[name + ".frame_data"] = np.frombuffer(bytes(fdata), dtype=np.uint8).copy()
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **out)
    print("wrote camera_poses.json (%d poses) and frames.npz (%d arrays)" % (len(poses), len(out)))


if __name__ == "__main__":
    main()
