"""Every `file:line` citation of the reference in the boundary header, the oracle and the design documents names a
file that exists in the reference tree and lines that exist in it (skipped where /root/reference is absent, e.g.
on the GPU box)."""
import os
import re

import pytest

from conftest import ROOT

REF = "/root/reference"
FILES = ["include/vkrt.h", "DESIGN.md", "INTEGRATION.md", "oracle/vkrt_oracle.cpp", "oracle/vkrt_oracle.h", "oracle/spirv_interp.py",
         "tests/test_spirv_pins.py", "tests/golden/make_spirv_vectors.py",
         "vk-renderer_b200/csrc/vkrt_api.cu", "vk-renderer_b200/csrc/vkrt_device.cuh", "vk-renderer_b200/csrc/vkrt_render.cu",
         "vk-renderer_b200/csrc/vkrt_wavefront.cu", "vk-renderer_b200/csrc/host/GraphicsDevice_cuda.cpp",
         "vk-renderer_b200/csrc/host/Camera.cpp", "vk-renderer_b200/csrc/host/headless_main.cpp", "vk-renderer_b200/device.py"]
KNOWN = {"Tracer.comp": "Assets/Tracer.comp", "Raytracer.comp": "Assets/Raytracer.comp", "Fullscreen.frag": "Assets/Fullscreen.frag",
         "Fullscreen.vert": "Assets/Fullscreen.vert", "GraphicsDevice.cpp": "Source/GraphicsDevice.cpp", "Main.cpp": "Source/Main.cpp",
         "Camera.cpp": "Source/Camera.cpp", "GraphicsDevice.h": "Include/GraphicsDevice.h", "Camera.h": "Include/Camera.h",
         "VulkanState.h": "Source/VulkanState.h"}
CITE = re.compile(r"\b((?:Assets|Source|Include)/)?([A-Za-z]+\.(?:comp|frag|vert|cpp|h)):(\d+)(?:-(\d+))?")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not mounted here")
def test_reference_citations_exist():
    lengths, n = {}, 0
    for rel in FILES:
        text = open(os.path.join(ROOT, rel), errors="replace").read()
        for m in CITE.finditer(text):
            name = m.group(2)
            if name not in KNOWN:
                continue                      # e.g. this repo's own files
            path = os.path.join(REF, KNOWN[name])
            if path not in lengths:
                assert os.path.exists(path), "%s cites %s, which the reference does not have" % (rel, KNOWN[name])
                lengths[path] = sum(1 for _ in open(path, errors="replace"))
            lo = int(m.group(3)); hi = int(m.group(4) or lo)
            assert 1 <= lo <= hi <= lengths[path], "%s cites %s:%s beyond its %d lines" % (rel, name, m.group(0).split(":")[1], lengths[path])
            n += 1
    assert n > 150, "only %d citations found: the pattern is broken" % n
