#!/usr/bin/env python
"""bench.py — headline benchmark of the render path: Mrays/s (and frame time) of one frame of a BASELINE.json config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1: one rank per GPU)

A "step" is one frame of the workload.  1 ray = one TraverseBVH-equivalent query (primary, continuation or shadow ray),
counted on the device and identical to the CPU oracle's count in parity mode.  Default workload = C4 (BASELINE.json
configs[3], the config the >= 1 Grays/s target is quoted on: synthetic 1 000 000-triangle height field, GPU-built LBVH,
3840x2160, depth 6, 1 spp).  With N > 1 the SAME frame is sharded by bands of `--band-rows` rows (default 8) over the ranks
(strong scaling) and gathered on rank 0 over NVLink: by default each rank's resolve kernel stores straight into rank 0's
frame through a CUDA-IPC peer mapping (no collective), `--gather nccl` uses torch.distributed.gather of packed bands instead.

Printed JSON (rank 0, ONE line on stdout; everything else goes to stderr): the driver's contract plus `roofline` (dominant kernel
family k_traverse, algorithmic bytes per SURVEY.md §8d ÷ CUDA-event time), `cpu_baseline` (the CPU oracle on a bounded sample of
the same frame) and `e2e` (the same metric through the host API with HOST output buffers, readback inside the timed region:
N = 1 rtb_render_begin / rtb_render_end with as many frames in flight as lanes; N > 1 five frame buffers on rank 0, three frames
in flight per rank, one barrier per frame, rank 0 reads every frame back on a host thread).
`--impl reference` times the CPU restatement of the reference renderer (oracle/) on the host cores: the reference itself
is Unity C# + HLSL and cannot run here.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, scene factory name, args, width, height, depth, spp)
    "c2": ("C2: reference sample scene test_scene_1 (1426 triangles), 1920x1080, depth 6, 1 spp", "sample", ("test_scene_1",), 1920, 1080, 6, 1),
    "c3": ("C3: 16x16 glass/mirror sphere grid + floor (196 620 triangles, tessellated like the reference), 3840x2160, depth 16, 1 spp",
           "spheres", (16,), 3840, 2160, 16, 1),
    "c4": ("C4: synthetic 1 000 000-triangle height field, GPU-built LBVH, 3840x2160, depth 6, 1 spp", "heightfield", (1000, 500), 3840, 2160, 6, 1),
    "c5": ("C5: C4 scene at 7680x4320, depth 6, 16 spp", "heightfield", (1000, 500), 7680, 4320, 6, 16),
}
# The oracle renders rows begin, begin+step, ... < end of the same frame: a bounded sample (about 10-30 s of CPU work).
# C3: the reference's median-split builder degenerates on this scene (3 nodes, two leaves of 98 310 triangles: its partition
# fails at the second level), so the CPU restatement needs ~30 s for ONE row; the sample is row 1080 alone.
CPU_ROWS = {"c2": (0, -1, 1), "c3": (1080, 1081, 1), "c4": (0, -1, 1), "c5": (0, -1, 32)}


def make_scene(kind, args):
    synth = importlib.import_module("cosig-raytracing_b200.synth")
    if kind == "sample":
        return synth.sample_scene(*args)
    if kind == "spheres":
        return synth.sphere_grid_scene(*args)
    return synth.heightfield_scene(*args)


def settings_for(w, h, depth, spp):
    scene_mod = importlib.import_module("cosig-raytracing_b200.scene")
    return scene_mod.RenderSettings(ResolutionOverride=(w, h), MaxDepth=depth, AASamples=spp)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would otherwise throttle the oracle)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_sample(workload, threads=0, analytic=False):
    threads = threads or host_threads()
    """The CPU oracle on every k-th row of the workload's frame.  Returns (counters, seconds, description)."""
    from oracle import oracle_py as O
    scene_mod = importlib.import_module("cosig-raytracing_b200.scene")
    desc, kind, args, w, h, depth, spp = WORKLOADS[workload]
    O.build()
    packed = scene_mod.pack_scene(make_scene(kind, args))
    t0 = time.time()
    osc = O.OracleScene.from_desc(packed.desc)
    if analytic:
        osc.set_primitive_mode(1)
    build_s = time.time() - t0
    b, e, step = CPU_ROWS[workload]
    if analytic and workload == "c3":
        b, e, step = 0, -1, 16  # no degenerate BVH in analytic mode: 257 primitives by brute force
    r = osc.render(settings_for(w, h, depth, spp).to_params(), rows=(b, e, step), threads=threads)
    c = r["counters"]
    rows = np.arange(b, h if e < 0 else e, step)
    return r, rows, c, build_s, (f"rows {b}:{h if e < 0 else e}:{step} of the {w}x{h} frame ({len(rows)} rows), same scene/settings; "
                                 f"BVH build {build_s:.2f} s not included")


def algorithmic_bytes_per_closest_ray(c):
    """32*n + 36*tau + 76*h with n, tau, h counted by the oracle on the reference-shape BVH (closest-hit rays)."""
    n_closest = c.rays_primary + c.rays_continuation
    n_bar = (c.nodes_visited - c.nodes_visited_shadow) / max(1, n_closest)
    tau_bar = (c.tris_tested - c.tris_tested_shadow) / max(1, n_closest)
    h = c.closest_hits / max(1, n_closest)
    return 32.0 * n_bar + 36.0 * tau_bar + 76.0 * h, dict(nodes_per_ray=round(n_bar, 3), tris_per_ray=round(tau_bar, 3), hit_fraction=round(h, 4))


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference renderer on the host cores (rank 0 only)."""
    if rank != 0:
        return
    desc, kind, sargs, w, h, depth, spp = WORKLOADS[args.workload]
    from oracle import oracle_py as O
    scene_mod = importlib.import_module("cosig-raytracing_b200.scene")
    O.build()
    packed = scene_mod.pack_scene(make_scene(kind, sargs))
    osc = O.OracleScene.from_desc(packed.desc)
    p = settings_for(w, h, depth, spp).to_params()
    b0, e0, step = CPU_ROWS[args.workload]
    step *= 2 if args.workload == "c4" else 1  # keep K+W steps within minutes
    rays = secs = 0.0
    threads = 0
    for i in range(args.warmup + args.steps):
        c = osc.render(p, rows=(b0 + (i % step if e0 < 0 else 0), e0, step), threads=host_threads())["counters"]
        if i >= args.warmup:
            rays += c.rays; secs += c.seconds
        threads = c.threads
    value = rays / secs / 1e6
    sample = f"each step = rows {b0}:{h if e0 < 0 else e0}:{step} of the {w}x{h} frame (offset rotates per step), all host threads"
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": desc, "note": "CPU restatement of the reference kernel (oracle/): the reference is Unity C# + HLSL and cannot run here"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  NCCL (its version banner), torch or the CUDA runtime may print to fd 1 as
    well, so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bvh", default="lbvh", choices=["lbvh", "reference"])
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--prim", default="tessellated", choices=["tessellated", "analytic"],
                    help="analytic: spheres / boxes are intersected analytically (SURVEY A13) instead of tessellated like the reference")
    ap.add_argument("--band-rows", type=int, default=8, help="N > 1: rows per band (multiple of 4); band b is rendered by rank b % N")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="diagnostic, N = 1 only: render just rank 0's bands of an N-way sharded frame (to profile one rank's share under ncu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--out-png", default=None, help="rank 0 writes the last frame here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if world >= 4:  # a rank's share of the frame is short: more lanes (streams with their own queues) overlap its per-depth tails (+5 % at N = 8)
        os.environ.setdefault("RTB_LANES", "6")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug lines must not share stdout with the JSON line
    import torch
    import torch.distributed as dist
    abi = importlib.import_module("cosig-raytracing_b200.abi")
    rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
    scene_mod = importlib.import_module("cosig-raytracing_b200.scene")
    bands = importlib.import_module("cosig-raytracing_b200.bands")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    desc, kind, sargs, w, h, depth, spp = WORKLOADS[args.workload]
    obj = make_scene(kind, sargs)
    st = settings_for(w, h, depth, spp)
    rt = rt_mod.RayTracer(devices=[local_rank], bvh_mode=abi.RTB_BVH_LBVH if args.bvh == "lbvh" else abi.RTB_BVH_REFERENCE,
                          primitive_mode=abi.RTB_PRIM_ANALYTIC if args.prim == "analytic" else abi.RTB_PRIM_TESSELLATED)
    packed = scene_mod.pack_scene(obj)
    stream = torch.cuda.ExternalStream(rt.stream(0), device=torch.device("cuda", local_rank))
    frame_bytes = w * h * 4

    # ---- where this rank's pixels go -----------------------------------------------------------------------------------
    p = st.to_params()
    BR = args.band_rows
    local = None
    if world == 1 and args.emulate_world > 1:
        p.band_rank, p.band_world, p.band_rows = 0, args.emulate_world, BR
        args.no_cpu_baseline = True
    if world > 1:
        p.band_rank, p.band_world, p.band_rows = rank, world, BR
    if world > 1 and args.gather == "peer":
        handle = [None]
        if rank == 0:
            ptr0, hbytes = rt.frame_export(frame_bytes)
            handle[0] = hbytes
        dist.broadcast_object_list(handle, src=0)
        dst_ptr = ptr0 if rank == 0 else rt.frame_import(handle[0])
        dst_bytes = frame_bytes
    elif world > 1:
        p.out_layout = abi.RTB_OUT_COMPACT
        local = torch.empty((max(1, bands.local_row_count(h, rank, world, BR)), w, 4), dtype=torch.uint8, device="cuda")
        dst_ptr, dst_bytes = local.data_ptr(), local.numel()
    else:
        dst_ptr, _ = rt.frame_export(frame_bytes)
        dst_bytes = frame_bytes

    def step_device(sync=False):
        rt.RenderToTexture(packed, p, dst_ptr, dst_bytes, sync=sync)
        if world > 1 and args.gather == "nccl":
            rt.synchronize()
            return bands.gather_bands(local[:bands.local_row_count(h, rank, world, BR)], h, w, rank, world, BR, dst=0)
        return None

    # ---- scene upload (first frame) --------------------------------------------------------------------------------------
    t0 = time.time()
    step_device(sync=True)
    first_frame_s = time.time() - t0
    s0 = rt.stats()
    for _ in range(args.warmup):
        step_device(sync=True)
    rays_rank = float(s0.rays_primary + s0.rays_continuation + s0.rays_shadow)
    rays_t = torch.tensor([rays_rank, float(s0.kernel_launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(rays_t)
    rays_frame, launches_frame = float(rays_t[0]), int(rays_t[1])

    # ---- device-timed steps: `value` ---------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    torch.cuda.synchronize(); rt.synchronize(); barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t_enq = time.perf_counter()
    for _ in range(args.steps):
        step_device(sync=False)
    enqueue_ms = (time.perf_counter() - t_enq) / args.steps * 1e3  # host time to enqueue one frame (launch-bound when ~ ms_per_step)
    rt.flush()  # the library alternates frames over two streams: make the timed stream wait for both
    ev1.record(stream)
    rt.synchronize(); torch.cuda.synchronize(); barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    rank_ms = [float(ms[0]) / args.steps]
    if world > 1:
        every = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(every, ms)
        rank_ms = [float(t[0]) / args.steps for t in every]  # per-rank device time: the spread is the load imbalance of the bands
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms[0]) / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer API: `e2e` ---------------------------------------------------------------------
    host = np.zeros((h, w, 4), np.uint8)
    lib = abi.load()
    import ctypes as C
    n_flight = max(2, min(8, int(os.environ.get("RTB_LANES", "4"))))  # frames in flight of the pipelined host API = the library's lanes
    pinned = [lib.rtb_alloc_pinned(frame_bytes) for _ in range(max(n_flight, 5))]
    host_views = [np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_uint8)), shape=(h, w, 4)) for pp in pinned]
    host_view = host_views[0]
    e2e_mode = f"rtb_render_begin/end, {n_flight} frames in flight, pinned host buffers" if world == 1 else "blocking per frame (barrier across ranks)"

    def run_e2e(n_steps):
        """N = 1: the pipelined host API — every step still copies its uniforms in and its RGBA8 frame out.  N > 1: blocking."""
        if world == 1:
            tickets = []
            for k in range(n_steps):
                if k >= n_flight:
                    rt.RenderEnd(tickets[k - n_flight])
                tickets.append(rt.RenderBegin(packed, st.to_params(), host_views[k % n_flight]))
            for t in tickets[-n_flight:]:
                rt.RenderEnd(t)
            return
        if args.gather == "nccl":
            for _ in range(n_steps):
                fr = step_device(sync=True)
                if rank == 0:
                    host_view[:] = fr.cpu().numpy()
            return
        # peer-store gather, software-pipelined over R contexts per rank: D frames are in flight on every rank; a frame's
        # bands land in frame buffer k % R of rank 0; once every rank has finished frame k-D (context sync + ONE barrier per
        # frame) a host thread of rank 0 reads it back while the following frames render.  The same barrier certifies that
        # rank 0 has finished reading frame k-R, whose buffer frame k overwrites.
        import threading
        readers = [None] * R

        def read_back(j):
            ring[j].frame_read(host_views[j])

        for k in range(n_steps + D):
            j = k % R
            if readers[j] is not None:
                readers[j].join()  # rank 0: frame k-R is in host memory
                readers[j] = None
            if k >= D:
                ring[(k - D) % R].synchronize()  # this rank's bands of frame k-D are stored
                barrier()
            if k < n_steps:
                ring[j].RenderToTexture(packed, p, ring_dst[j], frame_bytes, sync=False)
            if k >= D and rank == 0:
                j1 = (k - D) % R
                readers[j1] = threading.Thread(target=read_back, args=(j1,))
                readers[j1].start()
        for t in readers:
            if t is not None:
                t.join()
        host_view[:] = host_views[(n_steps - 1) % R]

    D, R = 3, 5  # frames in flight per rank, frame buffers (contexts) per rank
    ring, ring_dst = [rt] + [None] * (R - 1), [dst_ptr] + [None] * (R - 1)
    if world > 1 and args.gather == "peer":
        e2e_mode = (f"{R} contexts per rank, {D} frames in flight: frames k-{D - 1}..k render on all ranks while a host thread of rank 0 reads "
                    f"frame k-{D} back (one barrier per frame)")
        for j in range(1, R):
            ring[j] = rt_mod.RayTracer(devices=[local_rank], bvh_mode=rt.bvh_mode, primitive_mode=rt.primitive_mode)
            handle_j = [None]
            if rank == 0:
                ptr_j, hb = ring[j].frame_export(frame_bytes)
                handle_j[0] = hb
            dist.broadcast_object_list(handle_j, src=0)
            ring_dst[j] = ptr_j if rank == 0 else ring[j].frame_import(handle_j[0])
            ring[j].RenderToTexture(packed, p, ring_dst[j], frame_bytes, sync=True)  # uploads the scene to this context

    run_e2e(n_flight)
    torch.cuda.synchronize(); rt.synchronize(); barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize(); barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_s[0]) / args.steps * 1e3
    clocks = sampler.stop() if sampler else None
    # latency view of the same call: one blocking rtb_render per frame (no frames in flight)
    e2e_blocking = None
    if world == 1:
        nb = max(3, min(args.steps, 20))
        rt.RenderInto(packed, st.to_params(), host_view)
        t0 = time.perf_counter()
        for _ in range(nb):
            rt.RenderInto(packed, st.to_params(), host_view)
        bl = (time.perf_counter() - t0) / nb
        e2e_blocking = {"ms_per_frame": bl * 1e3, "value": rays_frame / bl / 1e6, "unit": "Mrays/s", "note": "rtb_render, one frame at a time"}
    host[:] = host_view

    # cold path: scene upload (H2D of the description + flatten + BVH build) + frame + readback, N = 1 only
    cold = None
    if world == 1:
        t0 = time.perf_counter()
        rt.InvalidateBVHCache()
        t1 = time.perf_counter()
        rt.RenderInto(packed, st.to_params(), host_view)
        cold_s = time.perf_counter() - t1
        sc = rt.stats()
        cold = {"ms": cold_s * 1e3, "ms_invalidate": (t1 - t0) * 1e3, "ms_upload_total": sc.ms_upload, "ms_build_device": sc.ms_build,
                "h2d_bytes": int(packed.desc.n_triangles) * 40, "note": "rtb_upload_scene + rtb_render, scene description in pageable host memory"}

    # ---- per-kernel-family times (profiling events per launch; separate, untimed pass) -------------------------------------
    rt.set_profiling(True)
    fam = np.zeros(3)
    n_prof = 3
    for _ in range(n_prof):
        step_device(sync=True)
        sp = rt.stats()
        fam += [sp.ms_traverse, sp.ms_shade, sp.ms_resolve]
    fam /= n_prof
    rt.set_profiling(False)
    sl = rt.stats()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        roof = {"bound": "hbm", "kernel": "k_traverse (every BVH query of the frame: closest-hit and shadow rays), all depths", "achieved": None,
                "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src}
        cpu = None
        if not args.no_cpu_baseline:
            ref, rows, c, build_s, sample = oracle_sample(args.workload, analytic=args.prim == "analytic")
            bpr, parts = algorithmic_bytes_per_closest_ray(c)
            closest_rays = float(s0.rays_primary + s0.rays_continuation)
            # k_traverse also serves the shadow rays: 32 n + 36 tau per shadow ray (no hit record is fetched)
            bps = (32.0 * c.nodes_visited_shadow + 36.0 * c.tris_tested_shadow) / max(1, c.rays_shadow)
            parts.update(shadow_nodes_per_ray=round(c.nodes_visited_shadow / max(1, c.rays_shadow), 3),
                         shadow_tris_per_ray=round(c.tris_tested_shadow / max(1, c.rays_shadow), 3), bytes_per_shadow_ray=round(bps, 1))
            achieved = (closest_rays * bpr + float(s0.rays_shadow) * bps) / (fam[0] * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(args.workload)
            # what the kernels actually requested: 64 B per LBVH node record (32 B per reference node), 48 B per triangle tested
            node_bytes = 64.0 if args.bvh == "lbvh" else 32.0
            requested = (float(s0.reserved[1]) * node_bytes + float(s0.reserved[2]) * 48.0) / (fam[0] * 1e-3) / 1e9
            if parts["tris_per_ray"] > 1000.0:  # the reference builder degenerated (C3): its counts say nothing about this traversal
                roof["note_degenerate_reference_bvh"] = ("the oracle's reference-shape BVH degenerates on this scene (leaves of ~10^5 triangles), so `achieved` "
                                                         "is computed from the bytes this traversal requested instead of the oracle's counts")
                achieved = requested
            roof.update(achieved=achieved, frac=achieved / peak, traffic=traffic, requested_gbs=requested, bytes_per_ray=round(bpr, 1), oracle_counts=parts,
                        launches_per_frame=depth + 1, ms_per_frame=float(fam[0]), rays_per_frame=closest_rays + float(s0.rays_shadow),
                        gpu_nodes_fetched_per_ray=round(s0.reserved[1] / max(1.0, rays_rank), 3), gpu_tris_tested_per_ray=round(s0.reserved[2] / max(1.0, rays_rank), 3),
                        note="algorithmic bytes = rays x (32 n + 36 tau + 76 h), n/tau/h counted by the CPU oracle on the reference-shape BVH "
                             "(SURVEY.md 8d); the LBVH visits fewer nodes than that, so frac can exceed what DRAM counters show")
            # per-GPU figures at N > 1: rank 0's rays over rank 0's kernel time; the CPU baseline itself is an N = 1 item
            cpu = {"value": c.rays / c.seconds / 1e6, "unit": "Mrays/s", "cores": int(c.threads), "kind": "port", "sample": sample,
                   "seconds": c.seconds} if world == 1 else None
            # the sampled rows must agree with the GPU frame (same scene, same settings): parity spot-check inside the bench
            d = np.abs(host[rows][..., :3].astype(np.int32) - ref["rgba8"][rows][..., :3].astype(np.int32)).max(axis=-1)
            parity = float((d <= 1).mean())
            if cpu is not None:
                cpu["parity_rows_within_1_255"] = parity
            roof["parity_rows_within_1_255"] = parity
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "bvh": args.bvh, "primitives": args.prim, "rays_per_frame": rays_frame, "frame_ms": ms_per_step,
                       "sharding": f"{world} rank(s), {BR}-row bands round-robin, gather={args.gather if world > 1 else 'none'}",
                       "rank_ms_per_step": [round(x, 4) for x in rank_ms], "host_enqueue_ms_per_frame": round(enqueue_ms, 4),
                       **({"emulated": f"rank 0 of {args.emulate_world} only (diagnostic run)"} if world == 1 and args.emulate_world > 1 else {}),
                       "l2": "no explicit flush: scene arrays (160 MB) plus ~1.4 GB of wavefront queues streamed per frame exceed the 126 MB L2",
                       "n_triangles": int(sl.n_triangles), "first_frame_s": first_frame_s},
            "clocks": clocks,
            "e2e": {"value": rays_frame / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms, "mode": e2e_mode,
                    "blocking": e2e_blocking,
                    "h2d_bytes_per_step": int(sl.h2d_bytes), "d2h_bytes_per_step": frame_bytes,
                    "note": "params -> pinned host RGBA8 frame through the C ABI; the scene stays resident like the reference's cached BVH "
                            "(RayTracer.cs:118-123); uniforms travel as kernel parameters"},
            "e2e_cold": cold,
            "gpu_launches": launches_frame * args.steps,
            "kernel_ms_per_frame": {"traverse": float(fam[0]), "shade": float(fam[1]), "resolve": float(fam[2])},
            "roofline": roof, "cpu_baseline": cpu,
        }
        emit(line)
        if args.out_png:
            from PIL import Image
            Image.fromarray(np.ascontiguousarray(host[::-1, :, :3])).save(args.out_png)
    for pp in pinned:
        lib.rtb_free_pinned(pp)
    barrier()
    for extra in ring[1:]:
        if extra is not None:
            extra.close()
    rt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
