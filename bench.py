#!/usr/bin/env python
"""bench.py -- headline benchmark: paths/s (and Mrays/s) of the path-tracing hot path on the Book-2 final
scene (BASELINE.json config 4: 800x800, max_depth 40), on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c4] [--spp-per-step S]

A "step" is one pass of the hot path over one batch: `spp-per-step` samples of EVERY pixel of the full-size
config (default 500; 20 steps = the config's 10000 spp = 6.4 G paths).  Throughput is independent of the
number of steps.  With N ranks every rank renders its own sample slices (weak scaling: per-GPU work fixed)
and the W*H float4 accumulators are summed with NCCL after each step, inside the timed region.

value     : paths/s over all ranks, scene resident in HBM, device time (CUDA events), max over ranks.
e2e       : the same metric through the C ABI with HOST buffers: gs_scene_create (scene H2D) + render +
            accumulator reduce + gs_resolve (image D2H) every step, wall clock bracketed by synchronizes.
roofline  : HBM-bound model; achieved = rays/s x algorithmic bytes per ray (bench_constants.json, frozen
            from the oracle's reference-semantics counters: 32 B per box test, 32 B per primitive test,
            16 B per texel), peak = MEASURED_PEAKS.json hbm_gbs.
--impl reference : the CPU restatement of the reference (oracle/, C++ port of the Rust) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line.  Libraries (NCCL prints its version banner there) get stderr instead.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


FULL_SPP = {"c1": 100, "c2": 1000, "c3": 1000, "c4": 10000, "c5": 4096}
WORKLOAD = {
    "c1": "Book-1 random spheres (config 1), 400x225, max_depth 50",
    "c2": "earthmap sphere + airport.hdr (config 2), 1280x720, max_depth 50",
    "c3": "Cornell box (config 3), 600x600, max_depth 50",
    "c4": "Book-2 final scene (config 4): BVH over 2400 box quads + 1005 spheres, 2 constant media, Perlin + image textures, 800x800, max_depth 40",
    "c5": "1M random spheres (config 5), 3840x2160, max_depth 50",
}


def constants():
    with open(os.path.join(ROOT, "bench_constants.json")) as f:
        return json.load(f)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(ln) for ln in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)  # gone before the end-to-end leg starts: its queries take the driver's lock
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(config, target_seconds=12.0):
    """The oracle (C++ port of the Rust reference) on all host cores, bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi as orc

    cores = os.cpu_count() or 1
    probe = orc.OracleScene.build(config, spp=1)
    _, _, c, sec = probe.render(seed=1, rng=orc.OracleScene.WYRAND, threads=cores)
    rate = c["paths"] / sec
    spp = max(1, min(64, int(target_seconds * rate / (probe.width * probe.height))))
    s = orc.OracleScene.build(config, spp=spp)
    _, _, c, sec = s.render(seed=2, rng=orc.OracleScene.WYRAND, threads=cores)
    out = {"value": c["paths"] / sec, "unit": "paths/s", "cores": cores, "kind": "port", "mrays_per_s": c["rays"] / sec / 1e6,
           "sample": f"{config} at full resolution {s.width}x{s.height}, {spp} spp ({c['paths']} paths, {sec:.1f} s), C++ port of the Rust reference, "
                     f"{cores} threads, Wyrand RNG"}
    # How the baseline scales with threads (1 spp each: a few seconds in total), and the same port with the per-hit material
    # reference count removed (the reference clones an Arc<dyn Material> per accepted hit, sphere.rs:102 / quad.rs:105: an atomic
    # counter every thread hammers when 2400 quads share one material) -- so the GPU/CPU ratio is stated against a baseline whose
    # behaviour is understood.  `value` above stays the faithful port on all cores.
    try:
        curve, t = {}, 1
        while t <= cores:
            _, _, c1, s1 = probe.render(seed=3, rng=orc.OracleScene.WYRAND, threads=t)
            curve[str(t)] = round(c1["paths"] / s1)
            t *= 2
        out["thread_scaling_paths_per_s"] = curve
        norc = orc.OracleScene.build(config, spp=max(1, spp // 2), counters="norc")
        _, _, c2, s2 = norc.render(seed=2, rng=orc.OracleScene.WYRAND, threads=cores)
        out["no_refcount"] = {"value": c2["paths"] / s2, "unit": "paths/s", "cores": cores,
                              "what": "same port, material handle = plain pointer instead of a ref-counted clone per accepted hit"}
    except Exception as e:  # a reported extra, never a reason to lose the bench line
        out["thread_scaling_error"] = str(e)[:200]
    return out


def run_reference(args):
    """--impl reference: times the CPU port on this box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi as orc

    cores = os.cpu_count() or 1
    probe = orc.OracleScene.build(args.config, spp=1)
    _, _, c, sec = probe.render(seed=1, rng=orc.OracleScene.WYRAND, threads=cores)
    rate = c["paths"] / sec
    total = max(1, args.steps + args.warmup)
    per_step_s = min(6.0, 150.0 / total)
    spp = max(1, int(per_step_s * rate / (probe.width * probe.height)))
    s = orc.OracleScene.build(args.config, spp=spp)
    for w in range(args.warmup):
        s.render(seed=100 + w, rng=orc.OracleScene.WYRAND, threads=cores)
    paths = rays = 0
    secs = 0.0
    for k in range(args.steps):
        _, _, c, sec = s.render(seed=200 + k, rng=orc.OracleScene.WYRAND, threads=cores)
        paths += c["paths"]; rays += c["rays"]; secs += sec
    value = paths / secs
    line = {"impl": "reference", "metric": "paths/s", "value": value, "unit": "paths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "mrays_per_s": rays / secs / 1e6,
            "config": {"workload": WORKLOAD[args.config], "samples_per_pixel_per_step": spp, "scene_seed": 1},
            "cpu_baseline": {"value": value, "unit": "paths/s", "cores": cores, "kind": "port",
                             "sample": f"{args.config} full resolution {s.width}x{s.height}, {spp} spp per step, {args.steps} steps; C++ port of the Rust "
                                       f"reference (the Rust itself cannot be built: no cargo/rustc), {cores} threads"},
            "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0



def issue_roofline(config):
    """roofline.issue: what the SM's issue slots and lanes did in the dominant kernel, from the committed ncu page of this round
    (bench_constants.json carries the numbers and names the file under profiles/ they come from).  The HBM line above is the
    contract's normalisation; for a kernel whose data lives in L1/L2 THIS is the bound that describes it."""
    k = constants().get(config, {})
    if "ncu_issue" not in k:
        return {}
    return {"issue": k["ncu_issue"]}


def run_extras(args, gs, gdist, dist, torch, np, scene, host, rank, local_rank, world, device, stream, barrier):
    """Three more measurements in the same JSON line (so the driver's 1/2/4/8-GPU runs carry them):
    c5      config 5 (10^6 spheres, 3840x2160): 16 spp per step per GPU, accumulator reduce inside the timed region
    strong  config 4 with a FIXED total of 4000 spp per step split over the ranks (strong scaling)
    single_process  (N > 1) the same weak-scaling step through gs_render_multi: rank 0 alone drives all N GPUs in one call"""
    import ctypes as C

    from grayshift_b200 import ffi

    extra = {}
    W, H = scene.width, scene.height

    def timed_steps(render_step, reduce_step, steps, warm=1):
        for w in range(warm):
            render_step(w); reduce_step()
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for k in range(steps):
            ev[k][0].record(stream); render_step(warm + k); ev[k][1].record(stream); reduce_step(); ev[k][2].record(stream)
        barrier()
        tot = torch.tensor([sum(e[0].elapsed_time(e[2]) for e in ev), sum(e[1].elapsed_time(e[2]) for e in ev)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot[0]) / steps, float(tot[1]) / steps

    # ---- strong scaling on the headline config: the TOTAL per step is fixed ----
    if args.config == "c4":
        total_spp = 4000
        b, e = gdist.sample_range(total_spp, rank, world)
        acc = torch.zeros((W * H, 4), dtype=torch.float32, device=device)
        ms, red = timed_steps(lambda k: scene.render_accumulate(acc.data_ptr(), seed=11, sample_begin=k * total_spp + b, sample_count=e - b, stream=stream.cuda_stream),
                              lambda: gdist.reduce_accumulators(acc, dst=0) if world > 1 else None, steps=2)
        extra["strong"] = {"workload": "config 4, 800x800", "spp_total_per_step": total_spp, "ms_per_step": ms, "reduce_ms": red,
                           "paths_per_s": W * H * total_spp / (ms * 1e-3), "scaling": "strong"}
        del acc

    # ---- config 5 ----
    try:
        t0 = time.perf_counter()
        h5 = gs.HostScene.build("c5", spp=16)
        h5.flatten()
        flat_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        s5 = gs.DeviceScene(h5, device=local_rank)
        torch.cuda.synchronize(device)
        create_ms = (time.perf_counter() - t0) * 1e3
        n5 = s5.width * s5.height
        acc5 = torch.zeros((n5, 4), dtype=torch.float32, device=device)
        s5.stats()
        ms, red = timed_steps(lambda k: s5.render_accumulate(acc5.data_ptr(), seed=1, sample_begin=(k * world + rank) * 16, sample_count=16, stream=stream.cuda_stream),
                              lambda: gdist.reduce_accumulators(acc5, dst=0) if world > 1 else None, steps=3)
        st5 = s5.stats()
        rays = torch.tensor([float(st5.rays), float(st5.paths)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(rays, op=dist.ReduceOp.SUM)
        k5 = constants()["c5"]
        paths_per_s = n5 * 16 * world / (ms * 1e-3)
        rays_per_path = float(rays[0]) / max(1.0, float(rays[1]))
        extra["c5"] = {"workload": WORKLOAD["c5"], "spp_per_step_per_gpu": 16, "ms_per_step": ms, "reduce_ms": red, "paths_per_s": paths_per_s,
                       "mrays_per_s": paths_per_s * rays_per_path / 1e6, "scaling": "weak", "bvh_width": s5.info("bvh_width"),
                       "scene_bytes": s5.info("scene_bytes"), "host_flatten_s": round(flat_s, 2), "scene_create_ms": round(create_ms, 1),
                       "accumulator_bytes": n5 * 16,
                       "hbm_model_frac": paths_per_s / world * rays_per_path * k5["bytes_per_ray"] / 1e9 / measured_peak()[0]}
        # end to end on rank 0's GPU alone: scene upload + render + resolve to pinned memory, every step
        if rank == 0:
            pin_rgb = torch.empty((n5, 3), dtype=torch.float32, pin_memory=True)
            pin_cnt = torch.empty(n5, dtype=torch.int32, pin_memory=True)
            np_rgb, np_cnt = pin_rgb.numpy(), pin_cnt.numpy().view("uint32")

            def e2e5(k):
                sc = gs.DeviceScene(h5, device=local_rank)
                a = torch.zeros((n5, 4), dtype=torch.float32, device=device)
                sc.render_accumulate(a.data_ptr(), seed=1, sample_begin=1000 + 16 * k, sample_count=16, stream=stream.cuda_stream)
                sc.resolve(a.data_ptr(), stream=stream.cuda_stream, out_rgb=np_rgb, out_counts=np_cnt)
                sc.close()
            e2e5(0)
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            for k in range(2):
                e2e5(1 + k)
            torch.cuda.synchronize(device)
            dt = (time.perf_counter() - t0) / 2
            extra["c5"]["e2e_single_gpu"] = {"paths_per_s": n5 * 16 / dt, "ms_per_step": dt * 1e3,
                                             "what": "gs_scene_create (10^6 spheres host -> device) + render + gs_resolve to pinned host memory + destroy, per step"}
        del acc5
        s5.close()
    except Exception as e:
        extra["c5"] = {"error": str(e)[:300]}

    # ---- the weak-scaling step again, single process: gs_render_multi drives every GPU from rank 0 ----
    if world > 1:
        cpu_pg = dist.new_group(backend="gloo")  # the other ranks wait on the CPU: no kernel of theirs spins on a GPU rank 0 uses
        torch.cuda.synchronize(device)
        dist.barrier(group=cpu_pg)
        if rank == 0:
            try:
                lib = ffi.cuda_lib()
                scenes = [scene] + [gs.DeviceScene(host, device=d) for d in range(1, world)]
                cam, st = scene.cam, ffi.gs_sample_settings()
                st.confidence, st.tolerance, st.batch_size, st.max_samples = 0.95, 0.0, 500 * world, 0
                rgb = torch.empty((W * H, 3), dtype=torch.float32, pin_memory=True)
                cnt = torch.empty(W * H, dtype=torch.int32, pin_memory=True)
                stats = ffi.gs_stats()
                handles = (C.c_void_p * world)(*[s._h for s in scenes])
                times = []
                for k in range(3):
                    t0 = time.perf_counter()
                    ffi.check(lib.gs_render_multi(handles, world, C.byref(cam), C.byref(st), 5 + k, rgb.numpy().ctypes.data, cnt.numpy().ctypes.data, C.byref(stats)))
                    times.append(time.perf_counter() - t0)
                dt = min(times[1:])
                extra["single_process"] = {"what": "gs_render_multi: one host thread, one call, every GPU; cross-GPU sum fused into the resolve over peer-mapped memory",
                                           "spp_per_gpu": 500, "ms_per_call": dt * 1e3, "paths_per_s": W * H * 500 * world / dt,
                                           "render_ms_slowest_gpu": stats.kernel_ms, "exchange_ms": stats.exchange_ms}
                for s in scenes[1:]:
                    s.close()
            except Exception as e:
                extra["single_process"] = {"error": str(e)[:300]}
        dist.barrier(group=cpu_pg)
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="grayshift_b200")
    ap.add_argument("--config", default="c4", choices=sorted(WORKLOAD))
    ap.add_argument("--spp-per-step", type=int, default=0)
    ap.add_argument("--max-leaf", type=int, default=0, help="BVH leaf size limit handed to the host builder (0 = its default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-extras", action="store_true", help="skip the config-5, strong-scaling and single-process extras")
    ap.add_argument("--bvh-width", type=int, default=0, choices=[0, 2, 8], help="0 = the library's choice by scene size")
    ap.add_argument("--wavefront-sort", type=int, default=0, choices=[0, 1], help="wavefront pipeline: material-sorted shading queues")
    ap.add_argument("--pipeline", default="megakernel", choices=["megakernel", "wavefront"],
                    help="device pipeline (the megakernel is the default, chosen from ncu evidence; see DESIGN.md section 6)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import grayshift_b200 as gs
    from grayshift_b200 import dist as gdist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    spp_step = args.spp_per_step or {"c1": 100, "c2": 200, "c3": 250, "c4": 500, "c5": 16}[args.config]
    host = gs.HostScene.build(args.config, spp=spp_step)
    t0 = time.perf_counter()
    host.flatten(max_leaf_size=args.max_leaf)
    flatten_s = time.perf_counter() - t0
    scene = gs.DeviceScene(host, device=local_rank)
    scene.set_option("wavefront", 1 if args.pipeline == "wavefront" else 0)
    scene.set_option("wavefront_sort", args.wavefront_sort)
    if args.bvh_width:
        scene.set_option("bvh_width", args.bvh_width)
    W, H = scene.width, scene.height
    n_px = W * H
    accum = torch.zeros((n_px, 4), dtype=torch.float32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    stream = torch.cuda.current_stream(device)

    def step(k):
        begin = (k * world + rank) * spp_step  # disjoint global sample slices
        scene.render_accumulate(accum.data_ptr(), seed=1, sample_begin=begin, sample_count=spp_step, stream=stream.cuda_stream)
        if world > 1:
            gdist.reduce_accumulators(accum, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for w in range(args.warmup):
        step(w)
    barrier()
    scene.stats()  # reset device counters
    accum.zero_()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        starts[k].record(stream)
        step(args.warmup + k)
        ends[k].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop() if rank == 0 else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    st = scene.stats()
    counts = torch.tensor([float(st.paths), float(st.rays)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_s = float(total_ms.item()) * 1e-3
    paths, rays = float(counts[0].item()), float(counts[1].item())

    # ---- end to end through the C ABI with host buffers ----
    # Every step: gs_scene_create (the flattened scene arrays go host -> device), a zeroed accumulator, the render, the
    # cross-GPU reduce, gs_resolve into PINNED host memory, gs_scene_destroy.  One untimed step first (the first resolve
    # of a process pays one-off driver set-up); stage times are host-clock with a device synchronize after each stage.
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    h2d = d2h = 0
    pin_rgb = torch.empty((n_px, 3), dtype=torch.float32, pin_memory=True)
    pin_cnt = torch.empty(n_px, dtype=torch.int32, pin_memory=True)
    np_rgb, np_cnt = pin_rgb.numpy(), pin_cnt.numpy().view("uint32")
    stages = {"scene_create_h2d": 0.0, "render": 0.0, "reduce": 0.0, "resolve_d2h": 0.0, "destroy": 0.0}

    def e2e_step(k, timed):
        nonlocal h2d, d2h
        t = [time.perf_counter()]

        def mark():
            torch.cuda.synchronize(device)
            t.append(time.perf_counter())
        sc2 = gs.DeviceScene(host, device=local_rank)                 # gs_scene_create: scene arrays host -> device
        sc2.set_option("wavefront", 1 if args.pipeline == "wavefront" else 0)
        acc2 = torch.zeros((n_px, 4), dtype=torch.float32, device=device)
        mark()
        begin = ((args.warmup + args.steps + k) * world + rank) * spp_step
        sc2.render_accumulate(acc2.data_ptr(), seed=1, sample_begin=begin, sample_count=spp_step, stream=stream.cuda_stream)
        mark()
        if world > 1:
            gdist.reduce_accumulators(acc2, dst=0)
        mark()
        if rank == 0:
            sc2.resolve(acc2.data_ptr(), stream=stream.cuda_stream, out_rgb=np_rgb, out_counts=np_cnt)  # image device -> pinned host
            d2h = np_rgb.nbytes + np_cnt.nbytes
        mark()
        h2d = int(sc2.stats().h2d_bytes) + 512  # scene arrays + camera/settings structs
        sc2.close()
        mark()
        if timed:
            for name, a, b in zip(stages, t[:-1], t[1:]):
                stages[name] += (b - a) * 1e3 / e2e_steps

    e2e_step(-1, False)
    barrier()
    e0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k, True)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - e0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = (n_px * spp_step * world * e2e_steps) / float(e2e_s.item())

    # what the DEVICE actually fetches per ray (the HBM model above counts the reference's traversal): one short launch of the
    # counting variant of the kernel -> node visits per ray x node size (64 B binary, 96 B eight-wide)
    device_bytes = None
    if rank == 0 and args.pipeline == "megakernel":
        try:
            scene.set_option("count_nodes", 1)
            scene.stats()
            probe = torch.zeros((n_px, 4), dtype=torch.float32, device=device)
            scene.render_accumulate(probe.data_ptr(), seed=3, sample_begin=0, sample_count=max(1, min(8, spp_step)), stream=stream.cuda_stream)
            torch.cuda.synchronize(device)
            stp = scene.stats()
            node_bytes = 96 if scene.info("bvh_width") == 8 else 64
            device_bytes = {"node_visits_per_ray": stp.node_visits / max(1, stp.rays), "node_bytes": node_bytes,
                            "node_bytes_per_ray": stp.node_visits / max(1, stp.rays) * node_bytes,
                            "note": "fetched from L1/L2, not from HBM; leaf records (32 B sphere / 96 B quad per test) come on top"}
            scene.set_option("count_nodes", 0)
            del probe
        except Exception as e:
            device_bytes = {"error": str(e)[:200]}
    extra = {}
    if not args.no_extras:
        extra = run_extras(args, gs, gdist, dist, torch, np, scene, host, rank, local_rank, world, device, stream, barrier)

    if rank == 0:
        k = constants()[args.config]
        peak, peak_src = measured_peak()
        # dominant kernel = render_fixed_kernel, one launch per step per GPU; per-GPU figures for the roofline
        rays_per_launch = rays / (args.steps * world)
        launch_s = total_s / args.steps
        achieved = rays_per_launch * k["bytes_per_ray"] / launch_s / 1e9
        line = {
            "metric": "paths/s", "value": paths / total_s, "unit": "paths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 primitive intersection + f32 boxes/shading", "data": "synthetic",
            "mrays_per_s": rays / total_s / 1e6, "rays_per_path": rays / paths,
            "config": {"workload": WORKLOAD[args.config], "samples_per_pixel_per_step": spp_step, "full_config_spp": FULL_SPP[args.config],
                       "scene_seed": 1, "render_seed": 1, "l2": "flushed between timed steps (256 MiB write); scene working set itself is L2-resident",
                       "parallelism": f"spp-split x{world}, NCCL reduce of the fp32 accumulator each step" if world > 1 else "single GPU",
                       "bvh": f"binned SAH, <={args.max_leaf or 1} prims/leaf, traversed as the {'eight-wide compressed' if scene.info('bvh_width') == 8 else 'binary'} tree",
                       "pipeline": args.pipeline, "host_flatten_s": round(flatten_s, 3)},
            "e2e": {"value": e2e_value, "unit": "paths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "what": "gs_scene_create + gs_render_accumulate + reduce + gs_resolve to pinned host memory + gs_scene_destroy, per step",
                    "stage_ms_rank0": {k2: round(v, 3) for k2, v in stages.items()}},
            "gpu_launches": int(st.kernel_launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": k.get("ncu_dram_bytes_per_launch") if args.pipeline == "megakernel" else None,
                         "traffic_note": k.get("ncu_dram_note"),
                         "peak_source": peak_src, "kernel": "render_fixed_kernel" if args.pipeline == "megakernel" else "wf_traverse_kernel + wf_shade_kernel + wf_generate_kernel",
                         "algorithmic_bytes_per_ray": k["bytes_per_ray"],
                         "note": "algorithmic bytes = reference-semantics box/prim/texel counts x 32/32/16 B; the scene is L1/L2 resident, so this is a "
                                 "model of work, not DRAM traffic (see profiles/ for dram__bytes and L2 throughput)"},
            "clocks": clk, "wall_s_timed_region": wall,
        }
        line["roofline"].update(issue_roofline(args.config))
        if device_bytes:
            line["roofline"]["device"] = device_bytes
        if extra:
            line["extra"] = extra
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.config)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
