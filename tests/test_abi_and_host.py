"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports exactly what
include/vkrt.h declares (no compute calls here), error behaviour without a GPU, scene generators,
the shard layout, and the N>1 gather path with world_size 2 on gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def declared_functions():
    src = open(os.path.join(ROOT, "include", "vkrt.h")).read()
    return sorted(set(re.findall(r"VKRT_API\s+[\w\s\*]+?\b(vkrt_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vk):
    names = declared_functions()
    assert len(names) >= 35 and "vkrt_draw" in names and "vkrt_create" in names
    lib = C.CDLL(vk._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libvkrt_cuda.so does not export %s" % n
    assert sorted(vk._lib.SIGNATURES) == names                   # the binding covers the header, nothing else
    out = subprocess.run(["nm", "-D", "--defined-only", vk._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vkrt_\w+)", out)))
    assert exported == names                                      # and nothing undeclared leaks out
    assert vk._lib.load().vkrt_version().startswith(b"libvkrt_cuda")


def test_python_mirror_of_the_public_structs_matches_the_header(vk, tmp_path):
    """The ctypes structures of vk-renderer_b200/_lib.py are a hand-written mirror of include/vkrt.h: a C program
    compiled against the header prints every public struct's size and field offsets, and they must agree."""
    L = vk._lib
    pairs = {"vkrt_camera_data": L.CameraData, "vkrt_frame_data": L.FrameData, "vkrt_triangle": L.Triangle,
             "vkrt_material": L.Material, "vkrt_create_info": L.CreateInfo, "vkrt_counters": L.Counters,
             "vkrt_bvh_info": L.BvhInfo, "vkrt_external_image": L.ExternalImage, "vkrt_exchange_handle": L.ExchangeHandle}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "vkrt.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, "%s.%s" % (cname, fname)


def test_library_is_sm100a_only(vk):
    out = subprocess.run(["cuobjdump", "-lelf", vk._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert "sm_100a" in archs and not (archs - {"sm_100a", "sm_52"}), archs   # sm_52: cudart's own static stub


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under vk-renderer_b200/ may import, link or open it."""
    pkg = os.path.join(ROOT, "vk-renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r"^\s*(import|from)\s+oracle\b", r"vkrt_oracle", r"\borc_\w+\s*\(", r"#include.*oracle", r"_ref/libref"):
                    assert not re.search(pat, text, re.M), "%s matches %s" % (f, pat)


def test_create_argument_checks_need_no_device(vk):
    """vkrt_create validates its struct before it looks for a device: the same answers with and without a GPU."""
    L = vk._lib
    lib = L.load()

    def create(**kw):
        d = dict(struct_size=C.sizeof(L.CreateInfo), width=64, height=64, spp=1, max_depth=1, integrator=L.INTEGRATOR_PATH,
                 variant=L.VARIANT_MEGAKERNEL, frames_in_flight=2, device_id=0)
        d.update(kw)
        ctx = C.c_void_p()
        rc = lib.vkrt_create(C.byref(L.CreateInfo(**d)), C.byref(ctx))
        assert rc != L.SUCCESS or has_gpu()
        if rc == L.SUCCESS:
            lib.vkrt_destroy(ctx)
        return rc, lib.vkrt_last_error_string(None)

    assert lib.vkrt_create(None, None) == L.BAD_ARG
    for kw in (dict(struct_size=12), dict(width=0), dict(height=0), dict(width=65537), dict(width=65536, height=65536),
               dict(integrator=2), dict(variant=2), dict(max_depth=256), dict(max_depth=257), dict(spp=(1 << 24) + 1)):
        rc, msg = create(**kw)
        assert rc == L.BAD_ARG and msg.startswith(b"[app] - err :: "), (kw, rc, msg)
    assert b"2^30 pixels" in create(width=65536, height=32768)[1]
    # in-library multi-GPU (device_ids / n_devices): the argument checks also come before any device is touched
    ids8 = (C.c_int32 * 8)(0, 1, 2, 3, 4, 5, 6, 7)
    dup = (C.c_int32 * 8)(0, 1, 1, 3, 4, 5, 6, 7)
    for kw, what in ((dict(n_devices=9, device_ids=ids8), b"n_devices > 8"),
                     (dict(n_devices=3, device_ids=dup), b"distinct"),
                     (dict(n_devices=2, device_ids=ids8, tile_shard_count=2), b"shards the frame by itself"),
                     (dict(n_devices=2, device_ids=ids8, flags=L.FLAG_NO_RESOLVE), b"NO_RESOLVE")):
        rc, msg = create(**kw)
        assert rc == L.BAD_ARG and what in msg, (kw, rc, msg)


def test_only_the_allowed_places_use_the_oracle():
    """Outside tests/, only bench.py (its cpu_baseline / --impl reference leg) and __graft_entry__ (build + smoke) may
    import the oracle; tools/ and the product package never do."""
    users = []
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "build", "gpurun_out", "__pycache__", "tests", "oracle", "baseline", ".pytest_cache")]
        for f in files:
            if f.endswith((".py", ".sh")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+(oracle|spirv_interp)\b", text, re.M):
                    users.append(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert sorted(users) == ["__graft_entry__.py", "bench.py"], users
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # ... and inside bench.py only the CPU arm's function does
    assert bench.count("import oracle") == 1 and bench.split("import oracle")[0].rsplit("\ndef ", 1)[1].startswith("cpu_render(")


@pytest.mark.skipif(has_gpu(), reason="exercises the no-GPU failure path")
def test_create_fails_loudly_without_gpu(vk):
    with pytest.raises(vk.VkrtError) as e:
        vk.Renderer(64, 64)
    assert e.value.code == vk._lib.NO_SUITABLE_GPU
    assert "[app] - err ::" in str(e.value) and "no CPU fallback" in str(e.value)
    dev = vk.GraphicsDevice()
    assert dev.Construct(vk.GraphicsDevice.CreateInfo(None, 3, 2, 64, False)) == vk.GraphicsDevice.Error.NO_SUITABLE_GPU
    assert dev.Destruct() == vk.GraphicsDevice.Error.UNKNOWN


def test_error_enum_matches_reference(vk):
    E = vk.GraphicsDevice.Error                                    # Include/GraphicsDevice.h:46-52
    assert (E.SUCCESS, E.NO_SUITABLE_GPU, E.NO_SUITABLE_SURFACE, E.UNKNOWN) == (0, 1, 2, 3)
    info = vk.GraphicsDevice.CreateInfo(None, 3, 2, 1024, False)   # Main.cpp:105-115
    assert (info.swapchainSize, info.framesInFlight, info.raytrace_resolution, info.debug) == (3, 2, 1024, False)
    assert (info.width, info.height) == (1024, 1024) and info.swapchain_extent == (1024, 768)


def test_scene_generators_are_deterministic(vk):
    s = vk.scenes
    assert s.tracer_default().digest() == s.tracer_default().digest()
    a, b = s.random_spheres(1024), s.random_spheres(1024)
    assert a.digest() == b.digest() == "ed102bc746ea584f"
    assert a.spheres.shape == (1025, 4) and a.materials.shape == (8 + 1024, 12)
    assert a.spheres[1:, 3].min() >= 0.5 and a.spheres[1:, 3].max() <= 3.0
    kinds = a.materials[8:].view(np.uint32)[:, 8]
    frac_diel = (kinds == 1).mean()
    metal = (a.materials[8:, 7] == 1.0).mean()
    assert 0.05 < frac_diel < 0.16 and 0.22 < metal < 0.38            # ~10 % dielectric, ~30 % metal, rest lambertian
    g = s.grid_spheres()
    assert g.spheres.shape == (100001, 4) and g.digest() == "d78035e0463ed3f3"
    c, r = g.spheres[1:, :3], g.spheres[1:, 3]
    assert c[:, 0].min() - r[0] >= -60 and c[:, 0].max() + r[0] <= 60 and c[:, 1].min() - r[0] >= 2 and c[:, 1].max() + r[0] <= 120
    assert s.random_spheres(64, seed=1).digest() != s.random_spheres(64, seed=2).digest()


def test_default_scenes_match_the_oracles_builtin_copy(vk, oracle):
    """scenes.py, libvkrt_cuda's vkrt_use_default_scene and the oracle each restate the shader constants;
    the python and oracle copies must render identically."""
    from helpers import apply_scene, bits_equal
    fd = vk.default_frame_data(aspect_ratio=4 / 3, seed=0.5)
    for scene, which, integ, depth in ((vk.scenes.tracer_default(), oracle.SCENE_TRACER, oracle.PATH, 4),
                                       (vk.scenes.raytracer_default(), oracle.SCENE_RAYTRACER, oracle.WHITTED, 2)):
        a, ia, _, _ = apply_scene(oracle, scene).render(fd, 48, 36, spp=2, max_depth=depth, integrator=integ, seed=3)
        b, ib, _, _ = oracle.Scene().use_default(which).render(fd, 48, 36, spp=2, max_depth=depth, integrator=integ, seed=3)
        assert bits_equal(a, b) and np.array_equal(ia, ib)


def test_shard_layout(vk):
    from vk_renderer_b200.sharding import owned_pixels, shard_layout
    assert shard_layout(0, 1) == ((0, 1), (0, 1))
    assert shard_layout(5, 8) == ((5, 8), (0, 1))
    assert shard_layout(5, 8, sample_shards=2) == ((1, 4), (1, 2))
    w, h = 200, 150
    seen = np.zeros(w * h, dtype=np.int32)
    for rank in range(3):
        pix = owned_pixels(w, h, rank, 3)
        valid = pix[pix >= 0]
        seen[valid] += 1
        assert pix.shape[0] % 1024 == 0
    assert np.all(seen == 1)                                       # tile shards partition the image exactly
    # the first 32 slots of a tile are an 8x4 pixel block
    p = owned_pixels(w, h, 0, 1)[:32]
    assert sorted((p % w).tolist()) == sorted(list(range(8)) * 4) and sorted((p // w).tolist()) == sorted(list(range(4)) * 8)


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import vk_renderer_b200 as V
import oracle as O
from vk_renderer_b200.sharding import owned_pixels, shard_layout, gather_packed
from helpers import bits_equal
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
w, h, spp = 100, 70, 4
fd = V.default_frame_data(aspect_ratio=w / h, seed=0.5)
sc = O.Scene().use_default(O.SCENE_TRACER)
# tile x sample layout: 2 ranks = 2 tile shards here; then 2 sample shards
for sample_shards in (1, 2):
    (tr, tc), (sr, scount) = shard_layout(rank, world, sample_shards)
    s0, s1 = sr * spp // scount, (sr + 1) * spp // scount
    full, _, _, _ = sc.render(fd, w, h, spp=spp, max_depth=4, seed=9, samples=(s0, s1), want_ids=False, want_rgba=False)
    pix = owned_pixels(w, h, tr, tc)
    n_floats = owned_pixels(w, h, 0, tc).shape[0] * 4
    send = torch.zeros(n_floats)
    flat = torch.from_numpy(full.reshape(-1, 4))
    valid = torch.from_numpy(pix >= 0)
    send.view(-1, 4)[: pix.shape[0]][valid] = flat[torch.from_numpy(pix[pix >= 0])]
    recv = gather_packed(send, rank, world)
    if rank == 0:
        acc = np.zeros((h * w, 4), dtype=np.float32)
        for src, buf in enumerate(recv):
            (t_r, t_c), (s_r, _) = shard_layout(src, world, sample_shards)
            p = owned_pixels(w, h, t_r, t_c)
            vals = buf.view(-1, 4)[: p.shape[0]].numpy()[p >= 0]
            if s_r > 0:
                acc[p[p >= 0]] = acc[p[p >= 0]] + vals
            else:
                acc[p[p >= 0]] = vals
        expect = None
        for s_r in range(sample_shards):
            expect, _, _, _ = sc.render(fd, w, h, spp=spp, max_depth=4, seed=9, samples=(s_r * spp // sample_shards, (s_r + 1) * spp // sample_shards),
                                        accum=expect, want_ids=False, want_rgba=False)
        assert bits_equal(acc.reshape(h, w, 4), expect), "gathered image differs (sample_shards=%%d)" %% sample_shards
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%%d_ok\n" %% rank); sys.stdout.flush()
"""


def test_two_rank_gather_on_gloo(tmp_path):
    """The N>1 path's host logic (tile ownership, packed layout, gather to rank 0, ordered sample-shard sum)
    with world_size 2 on CPU; the per-rank pixels come from the oracle instead of the GPU."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank0_ok" in r.stdout and "rank1_ok" in r.stdout


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference times the CPU restatement and prints the contract's JSON line."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "cfg2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    # under torchrun only rank 0 runs the CPU arm: the other ranks exit 0 without work or output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: the product arm of bench.py stops with a message instead of measuring something else."""
    if has_gpu():
        pytest.skip("exercises the no-GPU failure path")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout) and r.stdout.strip() == ""


# ---- the C++ host side (csrc/host): GraphicsDevice drop-in + headless frame loop ----------------------
def _headless():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_vkrt_build", os.path.join(ROOT, "vk-renderer_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    return b.build_host()


def test_cpp_camera_matches_reference_vector():
    """csrc/host/Camera.cpp (no glm) reproduces the reference's own Camera.cpp for Main.cpp's default view."""
    import json
    exe = _headless()
    out = subprocess.run([exe, "--print-camera"], capture_output=True, text=True, timeout=60).stdout.split()
    got = np.array([float.fromhex(v) for v in out], dtype=np.float32).reshape(4, 3)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "camera_poses.json")))["default_view"]["camera"]
    assert np.array_equal(got.view(np.uint32), np.array(gold, dtype=np.float32).view(np.uint32))


@pytest.mark.skipif(has_gpu(), reason="exercises the no-GPU failure path")
def test_cpp_host_reports_errors_like_the_reference():
    exe = _headless()
    r = subprocess.run([exe, "--frames", "2"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1
    assert "[app] - err :: no CUDA device" in r.stdout
    assert "[app] - err :: Graphics device creation failed :: 1" in r.stdout        # Main.cpp:121, Error::NO_SUITABLE_GPU
