#!/usr/bin/env python
"""Headline benchmark of the ray-tracing hot path (BASELINE.json: Mrays/s over all bounces, and
ms/frame at 1080p / 16 spp / depth 8).

  python bench.py [--gpus N --steps K --warmup W] [--workload cfg4|cfg2|cfg3] [--variant mega|wavefront]
  python bench.py --impl reference ...      # the CPU restatement of the reference shaders (oracle)

A "step" is one frame: one pass of the path tracer over every pixel of the workload, rendered from
scratch like the reference does every Draw (Source/GraphicsDevice.cpp:1215).  Default workload =
BASELINE.json configs[3], the configuration the north_star's target is quoted on: the synthetic
100,000-sphere scene at 1920x1080, 16 spp, depth 8, device LBVH.

Mrays/s counts trace_ray invocations of the reference algorithm (Tracer.comp:374): nearest-hit +
shadow queries, counted on the device by the kernels themselves.
N > 1 (torchrun, one rank per GPU), two shardings of the path:
  --scaling weak   (default) sample ranges: every GPU renders 16 spp of the whole frame, so N GPUs deliver a
                   16*N-spp frame per step (per-GPU work fixed); rank 0 sums the N accumulators in rank order
  --scaling strong interleaved 32x32 screen tiles of the one 16-spp frame (total work fixed)
Either way every step ends with the NCCL gather of the shards to rank 0 and the resolve there.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, width, height, spp, depth)
    "cfg4": ("synthetic 100k procedural spheres (jittered 50x40x50 grid + light + 5 planes), 1920x1080, 16 spp, depth 8, device LBVH",
             1920, 1080, 16, 8),
    "cfg3": ("synthetic 1,024 random spheres (mixed lambertian/metal/dielectric + light + 5 planes), 3840x2160, 64 spp, depth 8, device LBVH",
             3840, 2160, 64, 8),
    "cfg2": ("reference default scene (Tracer.comp:186-211), 1920x1080, 16 spp, depth 8, literal primitive loop",
             1920, 1080, 16, 8),
    # BASELINE.json configs[4]: the multi-GPU case (N > 1: interleaved tiles x 2 sample halves, --scaling strong implied);
    # one step is ~256 cfg4 frames of work, so run it with a small --steps
    "cfg5": ("synthetic 100k procedural spheres, 7680x4320, 256 spp, depth 8, device LBVH, tile x sample shards",
             7680, 4320, 256, 8),
}
SEED = 2026


def make_scene(V, workload):
    if workload in ("cfg4", "cfg5"):
        return V.scenes.grid_spheres(), True
    if workload == "cfg3":
        return V.scenes.random_spheres(1024), True
    return V.scenes.tracer_default(), False


def frame_data_for(V, w, h, step):
    # the reference passes a fresh seed = rand()/RAND_MAX every Draw (GraphicsDevice.cpp:1262); here a fixed sequence
    return V.default_frame_data(aspect_ratio=float(w) / float(h), seed=float((step * 0.61803398875) % 1.0))


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


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
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
    desc, w, h, spp, depth = WORKLOADS[workload]
    scene = {"cfg4": scenes.grid_spheres, "cfg5": scenes.grid_spheres, "cfg3": lambda: scenes.random_spheres(1024), "cfg2": scenes.tracer_default}[workload]()
    use_bvh = workload != "cfg2"
    sc = O.Scene(fast=True)
    sc.set_materials(scene.materials); sc.set_spheres(scene.spheres, scene.sphere_mat)
    sc.set_planes(scene.planes, scene.plane_mat); sc.set_triangles(scene.triangles, scene.tri_mat)
    if use_bvh:
        sc.build_bvh()
    mode = O.S_BVH if use_bvh else O.LITERAL
    cores = os.cpu_count() or 1

    def run(rect, step):
        fd = default_frame_data(aspect_ratio=float(w) / float(h), seed=float((step * 0.61803398875) % 1.0))
        t0 = time.perf_counter()
        _, _, _, cnt = sc.render(fd, w, h, spp=spp, max_depth=depth, integrator=O.PATH, sphere_mode=mode, seed=SEED,
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
            "ms_per_sample_step": 1e3 * sum(times) / len(times),
            "note": "CPU restatement of the reference shaders (oracle, -O3, std::thread over rows); lavapipe unavailable in image"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    desc, w, h, spp, depth = WORKLOADS[args.workload]
    per_step = max(2.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    r = cpu_render(args.workload, per_step, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "Mrays/s (all bounces)", "value": r["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_sample_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "spp": spp, "max_depth": depth},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": r["note"]}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import vk_renderer_b200 as V
    from vk_renderer_b200.sharding import FrameGather, shard_layout

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libvkrt_cuda has no CPU fallback (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    desc, w, h, spp, depth = WORKLOADS[args.workload]
    weak = world > 1 and args.scaling == "weak" and args.workload != "cfg5"
    if weak:
        spp *= world                  # every rank renders its own range of `spp / world` = the workload's samples
    # sample shards of the frame: weak = one per rank; cfg5 = tiles x 2 sample halves; else tiles only
    n_sample_shards = world if weak else (args.sample_shards or (2 if (args.workload == "cfg5" and world % 2 == 0) else 1))
    assert world % n_sample_shards == 0 and spp % n_sample_shards == 0
    scene, use_bvh = make_scene(V, args.workload)
    variant = V.VARIANT_WAVEFRONT if args.variant == "wavefront" else V.VARIANT_MEGAKERNEL
    tile_shard, sample_shard = shard_layout(rank, world, n_sample_shards)
    if args.shard_of > 1 and world == 1:          # diagnostic: one GPU renders tile shard 0 of N (no exchange)
        tile_shard = (0, args.shard_of)
    stream = torch.cuda.Stream(device=device)

    def make_renderer(flags):
        r = V.Renderer(w, h, spp=spp, max_depth=depth, integrator=V.INTEGRATOR_PATH, variant=variant, flags=flags,
                       device_id=local, tile_shard=tile_shard, sample_shard=sample_shard, stream=stream.cuda_stream)
        r.set_scene(scene)
        if use_bvh:
            r.build_bvh()
        r.set_seed(SEED)
        return r

    r = make_renderer(V.FLAG_NO_RESOLVE if world > 1 else 0)
    bvh = r.bvh_info()
    gather = FrameGather(r, rank, world, n_sample_shards, stream, device) if world > 1 else None
    n_px = w * h
    pinned = torch.empty(2, n_px * 4, dtype=torch.uint8).pin_memory() if rank == 0 else None

    def step(i, e2e=False):
        fd = frame_data_for(V, w, h, i)
        r.set_frame_index(i)
        r.draw(fd)
        if gather is not None:
            gather.gather()
            if rank == 0:
                r.resolve()
        if e2e and rank == 0:
            r.read_rgba8_async(pinned[i % 2].data_ptr(), n_px * 4)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # -------- device-resident timing ("value") ------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    r.reset_counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace_ms = []
    barrier()
    sampler.mark_begin()
    r.flush()
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    r.flush()                      # frames in flight end on the library's lane streams: order them before ev1
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    cnt = r.counters()
    launches_per_step = r.last_frame_timing()[2] + (1 if world > 1 else 0) + ((world + 1) if (world > 1 and rank == 0) else 0)
    rays = torch.tensor([cnt.closest_rays + cnt.shadow_rays, cnt.closest_rays, cnt.shadow_rays, cnt.paths,
                         cnt.shared_primary_rays, cnt.zero_term_shadow_rays], dtype=torch.float64, device=device)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    total_rays = float(rays[0].item())
    value = total_rays / (ms_total * 1e-3) / 1e6

    # duration of the dominant kernel(s), live: the library brackets every traversal-kernel launch (the megakernel,
    # or each extend / shadow launch of the wavefront) with CUDA events on the stream it launches on
    trav_ms, trav_launches = [], 0
    for i in range(min(args.steps, 5)):
        step(args.warmup + i)
        ms, nl, all_ms = r.last_frame_traversal_timing()
        trav_ms.append(ms); trav_launches = nl
        trace_ms.append(all_ms)
    barrier()
    kernel_ms = statistics.mean(trav_ms)
    frame_kernels_ms = statistics.mean(trace_ms)

    # the same frame with its waves one after the other on one stream (VKRT_FLAG_SERIAL_WAVES): per-launch event
    # times without the other lane's kernels sharing the SMs -- what the serialised ncu launch list can be compared with
    serial = None
    if variant == V.VARIANT_WAVEFRONT:
        try:
            rser = make_renderer(V.FLAG_NO_RESOLVE | V.FLAG_SERIAL_WAVES)
            s_trav, s_all = [], []
            for i in range(min(args.steps, 3) + 1):
                rser.set_frame_index(args.warmup + i)
                rser.draw(frame_data_for(V, w, h, args.warmup + i))
                ms, _, all_ms = rser.last_frame_traversal_timing()
                if i > 0:                                  # the first frame allocates the wave buffers
                    s_trav.append(ms); s_all.append(all_ms)
            rser.close()
            serial = (statistics.mean(s_trav), statistics.mean(s_all))
        except Exception as exc:                           # a measurement aid: never fails the bench
            sys.stderr.write("bench.py: serial-wave timing skipped: %r\n" % (exc,))
    barrier()

    # -------- end-to-end through the public API with host buffers ("e2e") -----------------------
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i, e2e=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = total_rays / float(e2e_t.item()) / 1e6

    # -------- algorithmic bytes of the dominant kernel (stats build, untimed) --------------------
    rs = make_renderer(V.FLAG_NO_RESOLVE | V.FLAG_STATS)
    n_stat = min(args.steps, 3)
    for i in range(n_stat):
        rs.set_frame_index(args.warmup + i)
        rs.draw(frame_data_for(V, w, h, args.warmup + i))
    sc = rs.counters()
    rs.close()
    traversed = (sc.closest_rays - sc.shared_primary_rays) + (sc.shadow_rays - sc.zero_term_shadow_rays)
    mega = variant == V.VARIANT_MEGAKERNEL
    # algorithmic bytes of the traversal kernels per frame: one BVH node per visit (the wavefront
    # walks the 32-byte quantised nodes, the megakernel the exact 64-byte ones) + 16 B per sphere fetched at a leaf +
    # per traversed ray its 32 B record read (origin, direction) and 8 B result write (wavefront) / per pixel the 16 B
    # accumulator store (megakernel, which also shades: + 56 B per nearest hit for sphere, material id and material)
    node_bytes = 64 if mega else 32
    alg_bytes = (sc.node_visits * node_bytes + sc.leaf_tests * 16) / n_stat
    if mega:
        alg_bytes += sc.closest_rays * (16 + 4 + 36) / n_stat + (n_px / world) * 16
    else:
        alg_bytes += traversed * (32 + 8) / n_stat
    hbm_peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = ncu_traffic(args.workload + ("_mega" if mega else "_wavefront"))
    roofline = {"bound": "hbm", "kernel": "k_path_mega<BVH>" if mega else "k_wf_trace<ANY,BVH> (extend + shadow launches of one frame)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src,
                "launches_per_frame": trav_launches, "kernel_ms_per_frame": kernel_ms,
                "kernel_ms_per_launch": kernel_ms / max(trav_launches, 1),
                "algorithmic_bytes_per_frame": alg_bytes, "algorithmic_bytes_per_launch": alg_bytes / max(trav_launches, 1),
                "share_of_step": kernel_ms / max(frame_kernels_ms, 1e-9),
                "share_note": "traversal launches / all kernel launches of a frame, both summed from per-launch CUDA events "
                              "(the wavefront's two lanes overlap, so the sums exceed ms_per_step)",
                "nodes_per_traversed_ray": sc.node_visits / max(traversed, 1),
                "leaf_tests_per_traversed_ray": sc.leaf_tests / max(traversed, 1),
                "node_bytes": node_bytes,
                "note": "traffic = ncu dram bytes per launch for the same launches (profiles/traffic.json). The tree (%.1f MB) is "
                        "L2/L1-resident, so achieved can exceed what DRAM delivers; the limiter is the SM's ALU pipe + lane "
                        "utilisation of the traversal loop (DESIGN.md section 4), the HBM fraction is the contract's reference number"
                        % (bvh.n_nodes * node_bytes / 1e6)}

    if serial is not None:
        roofline.update({"kernel_ms_per_frame_serial": serial[0], "achieved_serial": alg_bytes / (serial[0] * 1e-3) / 1e9,
                         "frac_serial": alg_bytes / (serial[0] * 1e-3) / 1e9 / hbm_peak,
                         "share_of_step_serial": serial[0] / max(serial[1], 1e-9),
                         "serial_note": "the same launches with the frame's two waves back to back on one stream "
                                        "(VKRT_FLAG_SERIAL_WAVES) instead of overlapping on two: each launch is timed alone, "
                                        "like in the serialised ncu launch list under profiles/; `achieved` / `frac` / "
                                        "`share_of_step` above come from the overlapped (shipping) configuration, where a "
                                        "launch shares the SMs with the other wave's kernels and its event time is longer"})
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_render(args.workload, 15.0)
        line = {"metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak" if (weak or world == 1) else "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "width": w, "height": h, "spp": spp, "max_depth": depth,
                           "variant": args.variant, "scene_sha": scene.digest(), "bvh_nodes": bvh.n_nodes,
                           "bvh_build_ms": bvh.build_ms, "parallelism": ("sample-shard x%d (%d spp per GPU, %d-spp frame)" % (world, spp // world, spp)) if weak
                           else ("tile-shard x%d" % world if n_sample_shards == 1 else
                                 "tile-shard x%d x sample-shard x%d (%d spp each)" % (world // n_sample_shards, n_sample_shards, spp // n_sample_shards)),
                           "l2_policy": "every step renders a new frame (new RNG keys); accumulator (%.0f MB) + rgba8 are rewritten each step; "
                                        "scene is L2-resident by design, no flush" % (n_px * 16 / 1e6)},
                "ms_per_frame": ms_total / args.steps,
                "rays_per_frame": total_rays / args.steps, "closest_rays": float(rays[1].item()) / args.steps,
                "shadow_rays": float(rays[2].item()) / args.steps, "paths_per_frame": float(rays[3].item()) / args.steps,
                "rays_note": "rays = trace_ray invocations of the reference algorithm; of these, shared_primary_rays (samples 2..S of a "
                             "pixel reuse the pixel's single primary-ray query) and zero_term_shadow_rays (unoccluded contribution exactly 0) "
                             "need no traversal of their own",
                "shared_primary_rays": float(rays[4].item()) / args.steps, "zero_term_shadow_rays": float(rays[5].item()) / args.steps,
                "traversed_rays_per_frame": (total_rays - float(rays[4].item()) - float(rays[5].item())) / args.steps,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 96 * world, "d2h_bytes_per_step": n_px * 4,
                        "ms_per_step": 1e3 * float(e2e_t.item()) / args.steps,
                        "note": "vkrt_draw(host FrameData) + resolve + rgba8 D2H into pinned memory each step, wall clock"},
                "gpu_launches": int(launches_per_step * args.steps),
                "clocks": clocks, "roofline": roofline}
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        band = tie_band(args.workload, scene.digest())
        if band is not None:
            line["config"]["primary_hit_tie_band"] = band
        if not use_bvh:
            # the compute-bound small scenes: algorithmic FLOP/s against the FP32 FFMA peak measured on this GPU
            # (~200 flops per brute-force trace_ray of the default Tracer scene, ~250 per DIFFUSE shading
            # event with its light sample = one per shadow ray here, ~60 per other shading event; FMA = 2 flops)
            fp32_peak = V.measure_fp32_peak(local)
            shading = float(rays[2].item())
            flops = total_rays * 200.0 + shading * 250.0 + max(float(rays[1].item()) - shading, 0.0) * 60.0
            line["fp32"] = {"achieved": flops / (ms_total * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                            "frac": flops / (ms_total * 1e-3) / 1e12 / fp32_peak, "peak_source": "measured FFMA chain (vkrt_measure_fp32_peak)",
                            "note": "algorithmic flops of the reference algorithm, not executed instructions; pipe utilisation "
                                    "from ncu is in profiles/ (metrics_k_path_mega_cfg2)"}
        if args.micro:
            line["fp32_peak_tflops_measured"] = V.measure_fp32_peak(local)
            line["l2_read_gbs_measured"] = V.measure_l2_bandwidth(local)
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="auto", choices=["auto", "mega", "wavefront"],
                    help="auto = wavefront for the LBVH scenes (cfg3/cfg4), megakernel for the 10-primitive default scene (cfg2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = sample-range shards (16 spp per GPU), strong = tile shards of the one 16-spp frame")
    ap.add_argument("--sample-shards", type=int, default=0,
                    help="N > 1, --scaling strong: split the ranks into tiles x this many sample ranges (default: 2 for cfg5, else 1)")
    ap.add_argument("--micro", action="store_true", help="also run the FP32 / L2 microbenchmarks")
    ap.add_argument("--shard-of", type=int, default=1, help="diagnostic (1 GPU): render only tile shard 0 of N, to size the fixed per-frame costs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.variant == "auto":
        args.variant = "mega" if args.workload == "cfg2" else "wavefront"
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
