#!/usr/bin/env python
"""Headline benchmark of the ray-tracing hot path (BASELINE.json: Mrays/s over all bounces, and
ms/frame at 1080p / 16 spp / depth 8).

  python bench.py [--gpus N --steps K --warmup W] [--workload cfg4|cfg1|cfg2|cfg2t|cfg3|cfg5]
  python bench.py --impl reference ...      # the CPU restatement of the reference shaders (oracle)

A "step" is one frame: one pass of the tracer over every pixel of the workload, rendered from
scratch like the reference does every Draw (Source/GraphicsDevice.cpp:1215).  Default workload =
BASELINE.json configs[3], the configuration the north_star's target is quoted on: the synthetic
100,000-sphere scene at 1920x1080, 16 spp, depth 8, device LBVH.

Mrays/s counts trace_ray invocations of the reference algorithm (Tracer.comp:374): nearest-hit +
shadow queries, counted on the device by the kernels themselves; `traversed_mrays_per_s` leaves out the
rays that need no traversal of their own (shared primary rays, zero-term shadow rays).

N > 1 (torchrun, one rank per GPU).  The headline `value` is STRONG scaling: the one 16-spp frame of the
workload is split into interleaved 32x32 screen tiles (total work fixed), every step ends with the frame
exchange to rank 0 and the resolve there.  The same line carries two more measurements as sub-objects:
  "weak"  sample-range shards: every GPU renders 16 spp of the whole frame (a 16*N-spp frame per step)
  "cfg5"  BASELINE.json configs[4]: 7680x4320, 256 spp, tiles x 2 sample halves
--exchange peer (default): every rank's last kernel stores its pixels straight into rank 0's memory over
NVLink (csrc/vkrt_exchange.cu); --exchange nccl: pack -> NCCL gather -> unpack.
At N = 1 the line also carries the other BASELINE configs as "workloads": cfg1, cfg2, cfg2t, cfg3, cfg5.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: description, size, sampling, integrator / scene / variant
WORKLOADS = {
    "cfg1": dict(desc="reference default scene of Raytracer.comp (:98-127), whitted integrator, 640x480, 1 spp, 2 bounces (BASELINE configs[0])",
                 w=640, h=480, spp=1, depth=2, integrator="whitted", scene="raytracer", variant="mega"),
    "cfg2": dict(desc="Raytracer.comp geometry (:98-127) through the path integrator: diffuse -> albedo, reflective -> metalness 1 / "
                      "roughness 0, else metalness 0 / roughness 0.4, + Tracer.comp's light sphere (SURVEY 8d); 1920x1080, 16 spp, depth 8, "
                      "64-frame progressive accumulation (BASELINE configs[1])",
                 w=1920, h=1080, spp=16, depth=8, integrator="path", scene="raytracer+light", variant="wavefront", progressive=64),
    "cfg2t": dict(desc="reference default scene of Tracer.comp (:186-211), 1920x1080, 16 spp, depth 8, literal primitive loop",
                  w=1920, h=1080, spp=16, depth=8, integrator="path", scene="tracer", variant="mega"),
    "cfg3": dict(desc="synthetic 1,024 random spheres (mixed lambertian/metal/dielectric + light + 5 planes), 3840x2160, 64 spp, depth 8, device LBVH (BASELINE configs[2])",
                 w=3840, h=2160, spp=64, depth=8, integrator="path", scene="random1024", variant="wavefront"),
    "cfg4": dict(desc="synthetic 100k procedural spheres (jittered 50x40x50 grid + light + 5 planes), 1920x1080, 16 spp, depth 8, device LBVH (BASELINE configs[3])",
                 w=1920, h=1080, spp=16, depth=8, integrator="path", scene="grid100k", variant="wavefront"),
    "cfg5": dict(desc="synthetic 100k procedural spheres, 7680x4320, 256 spp, depth 8, device LBVH, tile x sample shards (BASELINE configs[4])",
                 w=7680, h=4320, spp=256, depth=8, integrator="path", scene="grid100k", variant="wavefront"),
}
SEED = 2026


def make_scene(scenes, name):
    return {"grid100k": scenes.grid_spheres, "random1024": lambda: scenes.random_spheres(1024), "tracer": scenes.tracer_default,
            "raytracer": scenes.raytracer_default, "raytracer+light": lambda: scenes.raytracer_default(with_emitter=True)}[name]()


def uses_bvh(wl):
    return wl["scene"] in ("grid100k", "random1024")


def frame_data_for(default_frame_data, w, h, step):
    # the reference passes a fresh seed = rand()/RAND_MAX every Draw (GraphicsDevice.cpp:1262); here a fixed sequence
    return default_frame_data(aspect_ratio=float(w) / float(h), seed=float((step * 0.61803398875) % 1.0))


def parallelism(world, n_sample_shards, spp, weak):
    if world == 1:
        return "1 GPU"
    if weak:
        return "sample-shard x%d (%d spp per GPU, %d-spp frame)" % (world, spp // world, spp)
    if n_sample_shards == 1:
        return "tile-shard x%d" % world
    return "tile-shard x%d x sample-shard x%d (%d spp each)" % (world // n_sample_shards, n_sample_shards, spp // n_sample_shards)


def config_for(name, wl, variant, scene_sha, world, n_sample_shards, spp, weak):
    """The `config` object: identical for the product arm and the reference arm of the same command."""
    n_px = wl["w"] * wl["h"]
    return {"workload": wl["desc"], "name": name, "width": wl["w"], "height": wl["h"], "spp": spp, "max_depth": wl["depth"],
            "integrator": wl["integrator"], "variant": variant, "scene_sha": scene_sha,
            "parallelism": parallelism(world, n_sample_shards, spp, weak),
            "l2_policy": "every step renders a new frame (new RNG keys); accumulator (%.0f MB) + rgba8 are rewritten each step; "
                         "scene is L2-resident by design, no flush" % (n_px * 16 / 1e6)}


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line).  The sampler is started
    before the warm-up (nvidia-smi needs ~0.1-0.3 s to come up) and only the rows whose arrival time falls
    inside [mark_begin, mark_end] are used; if the timed region was shorter than the sampling period the
    rows taken under load during the warm-up are used and that is said in "window"."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= self.t1]
        window = "timed region"
        if not inside:
            inside = [r for t, r in self.rows if self.t0 is None or t <= self.t1]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        sm, mx, reasons = [], None, set()
        for r in inside:
            try:
                v = float(r[1]); mx = float(r[2])
            except Exception:
                continue
            sm.append(v)
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 500.0] or sm        # idle samples before the first launch read 120 MHz
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(busy), "window": window}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def tie_band(workload, scene_sha):
    """The north_star's "epsilon band of grazing ties whose count is reported": primary rays of the full-size frame for
    which the reference's literal in-order sphere loop names another sphere than rule S (counted by the CPU oracle when
    the primary-hit golden was made, tests/golden/make_cfg4_ids.py); outside it primary-hit ids are bit-exact."""
    if workload != "cfg4":
        return None
    try:
        import numpy as np
        z = np.load(os.path.join(ROOT, "tests", "golden", "cfg4_primary_ids.npz"))
        if str(z["scene_sha"][0]) != scene_sha:
            return None
        return {"primary_rays": int(z["tie_band"][0]), "of": int(z["ids"].size), "source": "tests/golden/cfg4_primary_ids.npz"}
    except Exception:
        return None


def ncu_evidence(key):
    """What the committed ncu capture of the dominant kernel says (profiles/ncu_limiters.json, written by
    tools/make_profiles_r2.py from ncu runs of this same command (tools/evidence_r2.sh)): DRAM bytes per launch and the units
    that bound the kernel.  Not measured by this run -- a profiler cannot run inside a timed bench -- so the entry
    carries the commit and file it came from."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_limiters.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference shaders; lavapipe / the shaders themselves cannot run
# in this image).  TEST INFRASTRUCTURE used here only as the reported baseline.
# ---------------------------------------------------------------------------------------------
def cpu_render(workload, target_seconds, steps=1, warmup=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import vk_renderer_b200.scenes as scenes   # scene generator only (numpy); no CUDA involved
    from vk_renderer_b200.device import default_frame_data
    O.build()
    wl = WORKLOADS[workload]
    w, h, spp, depth = wl["w"], wl["h"], wl["spp"], wl["depth"]
    scene = make_scene(scenes, wl["scene"])
    use_bvh = uses_bvh(wl)
    sc = O.Scene(fast=True)
    sc.set_materials(scene.materials); sc.set_spheres(scene.spheres, scene.sphere_mat)
    sc.set_planes(scene.planes, scene.plane_mat); sc.set_triangles(scene.triangles, scene.tri_mat)
    if use_bvh:
        sc.build_bvh()
    mode = O.S_BVH if use_bvh else O.LITERAL
    integrator = O.WHITTED if wl["integrator"] == "whitted" else O.PATH
    cores = os.cpu_count() or 1

    def run(rect, step):
        fd = frame_data_for(default_frame_data, w, h, step)
        t0 = time.perf_counter()
        _, _, _, cnt = sc.render(fd, w, h, spp=spp, max_depth=depth, integrator=integrator, sphere_mode=mode, seed=SEED,
                                 frame_index=step, rect=rect, want_ids=False, want_rgba=False)
        dt = time.perf_counter() - t0
        return cnt.closest_rays + cnt.shadow_rays, dt

    # probe a centred 1/64-area window to size the bounded sample
    pw, ph = max(w // 8, 8), max(h // 8, 8)
    px, py = (w - pw) // 2, (h - ph) // 2
    rays, dt = run((px, py, px + pw, py + ph), 0)
    rate = rays / dt
    frac = min(1.0, max(1.0 / 64.0, target_seconds * rate / (rays * 64.0)))
    sw, sh = max(int(w * frac ** 0.5) // 8 * 8, 8), max(int(h * frac ** 0.5) // 8 * 8, 8)
    sx, sy = (w - sw) // 2, (h - sh) // 2
    rect = (sx, sy, sx + sw, sy + sh)
    times, total_rays = [], 0
    for s in range(warmup + steps):
        rays, dt = run(rect, s)
        if s >= warmup:
            times.append(dt); total_rays += rays
    mrays = total_rays / sum(times) / 1e6
    sample = ("centred %dx%d window (%.1f%% of the pixels) of the %dx%d frame, same scene/spp/depth, %d step(s)"
              % (sw, sh, 100.0 * sw * sh / (w * h), w, h, steps))
    return {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample,
            "ms_per_sample_step": 1e3 * sum(times) / len(times), "scene_sha": scene.digest(),
            "note": "CPU restatement of the reference shaders (oracle, -O3, std::thread over rows); lavapipe unavailable in image"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    per_step = max(2.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    r = cpu_render(args.workload, per_step, steps=args.steps, warmup=args.warmup)
    world = args.gpus
    line = {"impl": "reference", "metric": "Mrays/s (all bounces)", "value": r["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_sample_step"],
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(args.workload, wl, args.variant, r["scene_sha"], world, 1, wl["spp"], False),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": r["note"]}
    band = tie_band(args.workload, r["scene_sha"])
    if band is not None:
        line["config"]["primary_hit_tie_band"] = band      # the same `config` object as the product arm prints
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
class Bench:
    """One process = one GPU.  measure() runs one workload in one sharding and returns its numbers (rank 0)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import vk_renderer_b200 as V
        self.torch, self.dist, self.V, self.args = torch, dist, V, args
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- libvkrt_cuda has no CPU fallback (use --impl reference for the CPU arm)")
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        assert self.world == args.gpus or self.world == 1, "launch with torchrun --nproc-per-node == --gpus"
        self.stream = torch.cuda.Stream(device=self.device)

    def barrier(self):
        self.torch.cuda.synchronize(self.device)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def allreduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return [float(x) for x in t.tolist()]

    def measure(self, name, steps, warmup, mode="strong", exchange="peer", variant=None, full=False, clocks=False, shard_of=1):
        """mode: "strong" (tile shards; cfg5: tiles x 2 sample halves) or "weak" (sample-range shards, spp * world).
        full: also the roofline of the dominant kernel, the e2e leg and the stats run."""
        torch, dist, V = self.torch, self.dist, self.V
        from vk_renderer_b200.sharding import FrameGather, PeerExchange, shard_layout
        rank, world, device, stream, local = self.rank, self.world, self.device, self.stream, self.local
        wl = WORKLOADS[name]
        w, h, spp, depth = wl["w"], wl["h"], wl["spp"], wl["depth"]
        variant = variant or wl["variant"]
        weak = world > 1 and mode == "weak"
        if weak:
            spp *= world                  # every rank renders its own range of `spp / world` = the workload's samples
        n_sample_shards = world if weak else (self.args.sample_shards or (2 if (name == "cfg5" and world % 2 == 0) else 1))
        assert world % n_sample_shards == 0 and spp % n_sample_shards == 0
        scene = make_scene(V.scenes, wl["scene"])
        use_bvh = uses_bvh(wl)
        whitted = wl["integrator"] == "whitted"
        progressive = int(wl.get("progressive", 0))
        if progressive:
            steps = progressive
        vflag = V.VARIANT_WAVEFRONT if variant == "wavefront" else V.VARIANT_MEGAKERNEL
        tile_shard, sample_shard = shard_layout(rank, world, n_sample_shards)
        if shard_of > 1 and world == 1:          # diagnostic: one GPU renders tile shard 0 of N (no exchange)
            tile_shard = (0, shard_of)

        def make_renderer(flags):
            r = V.Renderer(w, h, spp=spp, max_depth=depth, integrator=V.INTEGRATOR_WHITTED if whitted else V.INTEGRATOR_PATH,
                           variant=vflag, flags=flags, device_id=local, tile_shard=tile_shard, sample_shard=sample_shard,
                           stream=stream.cuda_stream)
            r.set_scene(scene)
            if use_bvh:
                r.build_bvh()
            r.set_seed(SEED)
            return r

        base_flags = V.FLAG_PROGRESSIVE if progressive else 0
        peer = world > 1 and exchange == "peer"
        r = make_renderer(base_flags | (V.FLAG_NO_RESOLVE if (world > 1 and (rank != 0 or not peer)) else 0))
        bvh = r.bvh_info()
        xch = PeerExchange(r, rank, world) if peer else None
        gather = FrameGather(r, rank, world, n_sample_shards, stream, device) if (world > 1 and not peer) else None
        n_px = w * h
        pinned = torch.empty(2, n_px * 4, dtype=torch.uint8).pin_memory() if (rank == 0 and full) else None

        def step(i, e2e=False):
            r.set_frame_index(i)
            r.draw(frame_data_for(V.default_frame_data, w, h, i))
            if gather is not None:
                gather.gather()
                if rank == 0:
                    r.resolve()
            if e2e and rank == 0:
                r.read_rgba8_async(pinned[i % 2].data_ptr(), n_px * 4)

        # -------- device-resident timing ("value") ------------------------------------------------
        sampler = ClockSampler(local) if (clocks and rank == 0) else None
        if sampler:
            sampler.start()
        for i in range(warmup):
            step(i)
        self.barrier()
        r.reset_counters()
        if progressive:
            r.reset_accum()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if sampler:
            sampler.mark_begin()
        r.flush()
        ev0.record(stream)
        for i in range(steps):
            step((0 if progressive else warmup) + i)       # progressive: frame index = 0..63 (seed = frame index)
        r.flush()                      # frames in flight end on the library's lane streams: order them before ev1
        ev1.record(stream)
        self.barrier()
        if sampler:
            sampler.mark_end()
        ms_total = ev0.elapsed_time(ev1)
        clocks_out = sampler.stop() if sampler else None
        cnt = r.counters()
        launches_per_step = r.last_frame_timing()[2]
        if gather is not None:
            launches_per_step += 1 + ((world + 1) if rank == 0 else 0)      # pack; unpack x world + resolve on rank 0
        rays = self.allreduce([cnt.closest_rays + cnt.shadow_rays, cnt.closest_rays, cnt.shadow_rays, cnt.paths,
                               cnt.shared_primary_rays, cnt.zero_term_shadow_rays], dist.ReduceOp.SUM)
        ms_total = self.allreduce([ms_total], dist.ReduceOp.MAX)[0]
        total_rays = rays[0]
        traversed = total_rays - rays[4] - rays[5]
        out = {"value": total_rays / (ms_total * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_total / steps,
               "traversed_mrays_per_s": traversed / (ms_total * 1e-3) / 1e6, "steps": steps, "warmup": warmup,
               "scaling": "weak" if weak else "strong", "n_gpus": world,
               "exchange": ("peer stores over NVLink (csrc/vkrt_exchange.cu)" if peer else "NCCL gather") if world > 1 else None,
               "config": config_for(name, wl, variant, scene.digest(), world, n_sample_shards, spp, weak),
               "bvh": {"nodes": bvh.n_nodes, "build_ms": bvh.build_ms, "lbvh_depth": bvh.depth,
                       "traversal_tree": "binned SAH, built on the device from the same leaves (exact union boxes)" if bvh.traversal_is_sah else "the LBVH",
                       "traversal_depth": bvh.traversal_depth, "traversal_build_ms": bvh.traversal_build_ms,
                       "note": "one-off per static scene, outside the timed region"} if use_bvh else None,
               "rays_per_frame": total_rays / steps, "closest_rays": rays[1] / steps, "shadow_rays": rays[2] / steps,
               "paths_per_frame": rays[3] / steps, "shared_primary_rays": rays[4] / steps,
               "zero_term_shadow_rays": rays[5] / steps, "traversed_rays_per_frame": traversed / steps,
               "gpu_launches": int(launches_per_step * steps), "clocks": clocks_out}
        if progressive:
            out["progressive_frames"] = progressive
            out["ms_per_accumulated_image"] = ms_total
        if shard_of > 1:
            out["config"]["parallelism"] = "diagnostic: tile shard 0 of %d on one GPU, no exchange" % shard_of

        if full:
            # -------- end-to-end through the public API with host buffers ("e2e") ---------------
            self.barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                step(warmup + i, e2e=True)
            self.barrier()
            e2e_s = self.allreduce([time.perf_counter() - t0], dist.ReduceOp.MAX)[0]
            out["e2e"] = {"value": total_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 96 * world,
                          "d2h_bytes_per_step": n_px * 4, "ms_per_step": 1e3 * e2e_s / steps,
                          "note": "vkrt_draw(host FrameData) + frame exchange + resolve + rgba8 D2H into pinned memory each step, wall clock"}
        if xch is not None:
            xch.close()
        r.close()

        if full and not whitted:
            # -------- the dominant kernel alone: per-launch CUDA events, waves serialised on one stream --------
            # (VKRT_FLAG_LAUNCH_TIMING | VKRT_FLAG_SERIAL_WAVES: the same launches as the shipping frame, bit-identical
            # image, but no other kernel shares the SMs while a launch is timed -- a kernel duration in the sense of the
            # serialised ncu launch list under profiles/; the shipping frame overlaps two waves and is FASTER than the sum)
            rt = make_renderer(V.FLAG_NO_RESOLVE | V.FLAG_SERIAL_WAVES | V.FLAG_LAUNCH_TIMING)
            t_trav, t_all, n_launch = [], [], 0
            for i in range(min(steps, 3) + 1):
                rt.set_frame_index(warmup + i)
                rt.draw(frame_data_for(V.default_frame_data, w, h, warmup + i))
                ms, n_launch, all_ms = rt.last_frame_traversal_timing()
                if i > 0:                                  # the first frame allocates the wave buffers
                    t_trav.append(ms); t_all.append(all_ms)
            rt.close()
            kernel_ms, frame_kernels_ms = statistics.mean(t_trav), statistics.mean(t_all)
            # -------- algorithmic bytes of the dominant kernel (stats build, untimed) --------------------
            rs = make_renderer(V.FLAG_NO_RESOLVE | V.FLAG_STATS)
            n_stat = min(steps, 3)
            for i in range(n_stat):
                rs.set_frame_index(warmup + i)
                rs.draw(frame_data_for(V.default_frame_data, w, h, warmup + i))
            sc = rs.counters()
            rs.close()
            trav = (sc.closest_rays - sc.shared_primary_rays) + (sc.shadow_rays - sc.zero_term_shadow_rays)
            mega = variant != "wavefront"
            # one BVH node per visit (wavefront: the compressed traversal nodes, megakernel: the exact
            # 64-byte ones) + 16 B per sphere fetched at a leaf + per traversed ray its 32 B record read and 8 B result
            # write (wavefront) / per pixel the 16 B accumulator store and 56 B of shading fetches per nearest hit (megakernel)
            node_bytes = 64 if mega else r_node_bytes(V)
            alg_bytes = (sc.node_visits * node_bytes + sc.leaf_tests * 16) / n_stat
            if mega:
                alg_bytes += sc.closest_rays * (16 + 4 + 36) / n_stat + (n_px / world) * 16
            else:
                alg_bytes += trav * (32 + 8) / n_stat
            hbm_peak, peak_src = measured_peaks()
            achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
            ev = ncu_evidence(name + ("_mega" if mega else "_wavefront")) or {}
            out["roofline"] = {
                "bound": "hbm", "kernel": "k_path_mega<BVH>" if mega else "k_wf_trace (the traversal launches of one frame)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": ev.get("dram_bytes_per_launch"), "traffic_source": ev.get("source", "no ncu capture committed for this workload"),
                "peak_source": peak_src, "launches_per_frame": n_launch, "kernel_ms_per_frame": kernel_ms,
                "kernel_ms_per_launch": kernel_ms / max(n_launch, 1),
                "algorithmic_bytes_per_frame": alg_bytes, "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                "share_of_step": kernel_ms / max(frame_kernels_ms, 1e-9),
                "timing": "each launch alone: per-launch CUDA events with the frame's waves back to back on one stream "
                          "(VKRT_FLAG_SERIAL_WAVES | VKRT_FLAG_LAUNCH_TIMING, bit-identical image); share_of_step = traversal launches / "
                          "all launches of that serialised frame (%.2f ms; the shipping frame overlaps its waves: ms_per_step)" % frame_kernels_ms,
                "nodes_per_traversed_ray": sc.node_visits / max(trav, 1), "leaf_tests_per_traversed_ray": sc.leaf_tests / max(trav, 1),
                "node_bytes": node_bytes,
                "bounding_unit": ev.get("limiters"),
                "note": "bound/frac are the contract's HBM figure: algorithmic bytes / kernel time against the measured copy bandwidth. "
                        "The tree (%.1f MB) is L1/L2-resident, so DRAM carries only the ray records (`traffic`); the units that actually "
                        "bound the kernel are listed in `bounding_unit` (from the committed ncu capture) and `l2`; that is also why "
                        "`frac` can exceed 1: the node bytes come out of L1 / L2, not HBM" % (bvh.n_nodes * node_bytes / 1e6)}
            if rank == 0:
                l2_peak = V.measure_l2_bandwidth(local)
                out["roofline"]["l2"] = {"achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                                         "peak_source": "measured: vkrt_measure_l2_bandwidth (L2-resident read loop, this GPU, this run)"}
        return out


def r_node_bytes(V):
    """Bytes of one traversal node of the wavefront trace kernel (what one node visit fetches)."""
    return int(getattr(V, "TRAVERSAL_NODE_BYTES", 32))


def fp32_object(V, local, m):
    """The compute-bound small scenes: algorithmic FLOP/s against the FP32 FFMA peak measured on this GPU (SURVEY.md 8d:
    ~200 flops per brute-force trace_ray of the default Tracer scene / ~150 of the Raytracer scene, ~250 per DIFFUSE shading
    event with its light sample = one per shadow ray here, ~60 per other shading event; FMA = 2 flops)."""
    fp32_peak = V.measure_fp32_peak(local)
    per_ray = 150.0 if "Raytracer.comp" in m["config"]["workload"] else 200.0
    frames = m["steps"]
    total, shading, closest = m["rays_per_frame"] * frames, m["shadow_rays"] * frames, m["closest_rays"] * frames
    flops = total * per_ray + shading * 250.0 + max(closest - shading, 0.0) * 60.0
    secs = m["ms_per_step"] * frames * 1e-3
    return {"achieved": flops / secs / 1e12, "peak": fp32_peak, "unit": "TFLOP/s", "frac": flops / secs / 1e12 / fp32_peak,
            "peak_source": "measured FFMA chain (vkrt_measure_fp32_peak, this GPU, this run)",
            "note": "algorithmic flops of the reference algorithm, not executed instructions; pipe utilisation from ncu is in profiles/"}


def run_ours(args):
    b = Bench(args)
    V, world, rank = b.V, b.world, b.rank
    name = args.workload
    wl = WORKLOADS[name]
    mode = args.scaling if world > 1 else "strong"
    if name == "cfg5":
        mode = "strong"
    m = b.measure(name, args.steps, args.warmup, mode=mode, exchange=args.exchange, variant=args.variant, full=True, clocks=True,
                  shard_of=args.shard_of)
    extras = {}
    if not args.no_extras and name == "cfg4" and args.shard_of == 1:
        if world > 1:
            other = "weak" if mode == "strong" else "strong"
            extras[other] = b.measure("cfg4", 5, 3, mode=other, exchange=args.exchange)
            extras["cfg5"] = b.measure("cfg5", 1, 1, mode="strong", exchange=args.exchange)
            extras["nccl_gather"] = b.measure("cfg4", 5, 3, mode=mode, exchange="nccl" if args.exchange == "peer" else "peer")
        else:
            w = {}
            w["cfg1"] = b.measure("cfg1", 50, 5)
            w["cfg2"] = b.measure("cfg2", 64, 3)
            w["cfg2t"] = b.measure("cfg2t", 10, 3)
            w["cfg3"] = b.measure("cfg3", 3, 3, full=True)
            w["cfg5"] = b.measure("cfg5", 1, 1)
            for k in ("cfg1", "cfg2", "cfg2t"):
                w[k]["fp32"] = fp32_object(V, b.local, w[k])
            extras["workloads"] = w
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_render(name, 15.0)
        line = {"metric": "Mrays/s (all bounces)", "value": m["value"], "unit": "Mrays/s", "n_gpus": world, "steps": m["steps"],
                "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
                "scaling": m["scaling"] if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": m["config"],
                "ms_per_frame": m["ms_per_step"], "traversed_mrays_per_s": m["traversed_mrays_per_s"],
                "rays_note": "rays = trace_ray invocations of the reference algorithm (value); of these, shared_primary_rays (samples 2..S of a "
                             "pixel reuse the pixel's single primary-ray query) and zero_term_shadow_rays (unoccluded contribution exactly 0) "
                             "need no traversal of their own: traversed_mrays_per_s counts only the rays that were traversed",
                "exchange": m["exchange"], "bvh": m["bvh"]}
        for k in ("rays_per_frame", "closest_rays", "shadow_rays", "paths_per_frame", "shared_primary_rays", "zero_term_shadow_rays",
                  "traversed_rays_per_frame", "e2e", "gpu_launches", "clocks", "roofline"):
            if k in m:
                line[k] = m[k]
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        band = tie_band(name, m["config"]["scene_sha"])
        if band is not None:
            line["config"]["primary_hit_tie_band"] = band
        if not uses_bvh(wl):
            line["fp32"] = fp32_object(V, b.local, m)
        if args.micro:
            line["fp32_peak_tflops_measured"] = V.measure_fp32_peak(b.local)
            line["l2_read_gbs_measured"] = V.measure_l2_bandwidth(b.local)
        for k, v in extras.items():
            if k == "workloads":
                for vv in v.values():
                    vv.pop("clocks", None)
            elif isinstance(v, dict):
                v.pop("clocks", None)
            line[k] = v
        if extras:
            line["extras_note"] = ("sub-objects are further measurements of the same command with fewer steps (their own steps/warmup "
                                   "are stated in each); cfg5 runs 1 warm-up + 1 timed frame")
        print(json.dumps(line))
    if world > 1:
        b.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="auto", choices=["auto", "mega", "wavefront"],
                    help="auto = wavefront for the LBVH scenes (cfg3/cfg4/cfg5), megakernel for the 10-primitive default scenes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline measurement (no weak / cfg5 / workloads sub-objects)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = tile shards of the one 16-spp frame, weak = sample-range shards (16 spp per GPU)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = stores into rank 0's memory over NVLink from the frame's last kernel; nccl = pack/gather/unpack")
    ap.add_argument("--sample-shards", type=int, default=0,
                    help="N > 1, strong: split the ranks into tiles x this many sample ranges (default: 2 for cfg5, else 1)")
    ap.add_argument("--micro", action="store_true", help="also run the FP32 / L2 microbenchmarks")
    ap.add_argument("--shard-of", type=int, default=1, help="diagnostic (1 GPU): render only tile shard 0 of N, to size the fixed per-frame costs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.variant == "auto":
        args.variant = WORKLOADS[args.workload]["variant"]
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
