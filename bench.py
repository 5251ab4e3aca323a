#!/usr/bin/env python3
"""bench.py — the BASELINE metric on the BASELINE configs.

Default workload (what the driver runs): BASELINE config 2, `scenes/caustics` at 1024x1024, 8 fixed spp, MAX_DEPTH 64, 1 M caustic
photons.  A step = one frame: the row loop of RayTracer::run (raytracer.h:93-160) over the whole image with the photon map already
built — the reference's own "second run()" timing — plus, reported beside it, the isolated photon gather over the frame's
primary-hit queries.  `--config C1..C5` selects another BASELINE config (C3-C5 render a stated number of samples per step).

  metric  Mrays/s (all bounces) = (closest-hit + shadow traversals issued by the frame) / frame time
  value   whole-job throughput with everything resident in HBM, CUDA events on the context's stream, max over ranks
  e2e     same metric through the host-pointer C ABI: gi_scene_upload + render + resolve with HOST buffers, ending in ONE host
          image on rank 0 (N > 1: the framebuffer collective is inside the timed region)
  gather  photon-gather Mqueries/s (gi_photon_gather_dev) over the primary-hit queries, same frame
  N > 1   --split samples (default; weak scaling): rank r renders samples [spp*r, spp*(r+1)) of every pixel, per-step fp64 reduce
          of the partial sums onto rank 0 (gi_framebuffer_reduce);
          --split tiles (strong scaling): the SAME frame, rank r renders the interleaved 16-row blocks r, r+N, .. in one wavefront
          (gi_render_rows), resolves them and rank 0 gathers the 8-bit rows (gi_framebuffer_gather) — all inside the timed region.
          The default line also carries `tile_split`: the strong-scaling numbers of the same frame at this N.
          The photon map is built once on rank 0 and broadcast as one slab (gi_photon_map_bcast), outside the timed region like the
          reference's cached map.  All collectives are the C ABI's own (NCCL from C); torch.distributed only carries the 128-byte id.

`--impl reference` times the reference's own OpenMP CPU path (oracle/_ref/gi_ref_fast, the unmodified reference sources) on a
bounded sample of the same workload, on the host cores of this box.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: scene, frame, samples per pixel of the config, samples rendered per bench step (per GPU under the sample split), MAX_DEPTH,
# caustic photons, the bounded CPU sample (rows of the same frame, full spp of the step)
CONFIGS = {
    "C1": dict(scene="cornell", w=512, h=512, spp=16, step_spp=16, depth=4, photons=750000, cpu_rows=(224, 288),
               note="reference assets, dragon.obj not mounted"),
    "C2": dict(scene="caustics", w=1024, h=1024, spp=8, step_spp=8, depth=64, photons=1000000, cpu_rows=(448, 512),
               note="reference assets, dragon.obj not mounted"),
    "C3": dict(scene="glass", w=1920, h=1080, spp=64, step_spp=16, depth=64, photons=275000, cpu_rows=(536, 544),
               note="reference assets, glass.obj not mounted; a step renders 16 of the 64 spp"),
    "C4": dict(scene="foliage", w=1920, h=1080, spp=256, step_spp=8, depth=64, photons=0, cpu_rows=(600, 604),
               note="seeded stand-in (scenes/make_standins.py): 12000 alpha-textured cards; a step renders 8 of the 256 spp"),
    "C5": dict(scene="sponza", w=3840, h=2160, spp=1024, step_spp=8, depth=64, photons=0, cpu_rows=(1080, 1082),
               note="seeded stand-in (scenes/make_standins.py): 262144-triangle atrium; a step renders 8 of the 1024 spp"),
}
BLOCK_ROWS = 16   # tile split: interleaved blocks of this many rows (the reference's row loop is `schedule(dynamic, 10)`)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def scene_file(cfg):
    return os.path.join(ROOT, "scenes", cfg["scene"], cfg["scene"] + ".scn")


def workload_name(name, cfg, spp):
    ph = f", {cfg['photons']} caustic photons, k=32 gather" if cfg["photons"] else ", no caustic photons"
    return f"{name} scenes/{cfg['scene']} {cfg['w']}x{cfg['h']}, {spp} spp per step (config: {cfg['spp']}), MAX_DEPTH {cfg['depth']}{ph}"


# ---- the reference's CPU path (oracle/_ref) -------------------------------------------------------------------------------
def run_reference_sample(name, cfg, repeat, threads=None):
    """Run gi_ref_fast (the unmodified reference, all host cores) on the bounded sample; per-repeat Mrays/s, gather Mq/s, core count."""
    exe = os.path.join(ROOT, "oracle", "_ref", "gi_ref_fast")
    if not os.path.exists(exe):
        raise FileNotFoundError(exe + " (build with `make -C oracle ref` where /root/reference is mounted)")
    threads = threads or os.cpu_count()
    out = tempfile.mkdtemp(prefix="gi_cpu_")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close")
    y0, y1 = cfg["cpu_rows"]
    spp = cfg["step_spp"]
    cmd = [exe, scene_file(cfg), out, "--w", str(cfg["w"]), "--h", str(cfg["h"]), "--y0", str(y0), "--y1", str(y1), "--s0", "0", "--s1", "1", "--samples", str(spp),
           "--max-depth", str(cfg["depth"]), "--photons", str(cfg["photons"]), "--repeat", str(repeat), "bench-frame"] + (["time-gather"] if cfg["photons"] else [])
    t0 = time.time()
    with open(os.path.join(out, "log.txt"), "w") as log:
        subprocess.check_call(cmd, stdout=log, stderr=subprocess.STDOUT, env=env, cwd=ROOT)
    wall = time.time() - t0
    meta = {}
    for line in open(os.path.join(out, "meta.txt")):
        if "=" in line:
            k, v = line.strip().split("=", 1)
            meta[k] = float(v)
    steps = []
    for r in range(repeat):
        rays = meta[f"bench_frame_trace_rays_{r}"] + meta[f"bench_frame_shadow_rays_{r}"]
        steps.append({"s": meta[f"bench_frame_s_{r}"], "rays": rays, "mrays": rays / meta[f"bench_frame_s_{r}"] / 1e6,
                      "gather_mq": (meta["time_gather_queries"] / meta[f"time_gather_s_{r}"] / 1e6) if cfg["photons"] else None})
    return {"steps": steps, "cores": threads, "wall_s": wall, "photon_s": meta.get("photon_trace_s", 0.0) + meta.get("photon_build_s", 0.0), "photons": meta.get("photons_stored", 0),
            "sample": f"rows {y0}-{y1 - 1} of the {cfg['w']}x{cfg['h']} frame ({(y1 - y0) * cfg['w']} pixels x {spp} spp, MAX_DEPTH {cfg['depth']}, {cfg['photons']} photons)"
                      + ("; gather: primary-hit queries of those rows" if cfg["photons"] else "")}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    K, Wm = args.steps, args.warmup
    try:
        res = run_reference_sample(args.config, cfg, K + Wm)
    except Exception as e:  # the oracle always exists in a built tree; report honestly if the binary is missing
        emit({"impl": "reference", "unavailable": str(e).splitlines()[0][:200]})
        return
    timed = res["steps"][Wm:]
    tot_rays = sum(s["rays"] for s in timed)
    tot_s = sum(s["s"] for s in timed)
    value = tot_rays / tot_s / 1e6
    gq = [s["gather_mq"] for s in timed if s["gather_mq"] is not None]
    line = {
        "impl": "reference", "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": K, "warmup": Wm,
        "ms_per_step": 1e3 * tot_s / K, "higher_is_better": True, "scaling": "weak" if args.split == "samples" else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "reference scene assets as mounted (synthetic stand-ins for C4/C5); bounded sample per step",
        "config": {"workload": workload_name(args.config, cfg, cfg["step_spp"]) + " (bounded sample per step)", "sample": res["sample"],
                   "impl": "oracle/_ref/gi_ref_fast: the unmodified reference sources (-O3 -march=x86-64-v3, OpenMP schedule(dynamic,10) row loop around RayTracer::radiance, no omp critical setPixel)"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                         "gather_mqueries_s": statistics.median(gq) if gq else None, "photon_phase_s": res["photon_s"]},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- our arm ---------------------------------------------------------------------------------------------------------------
def bytes_closest(rays, nodes, prims):   # B_trace = 64 + 64*N_node + 80*N_prim + 32 per ray
    return 96 * rays + 64 * nodes + 80 * prims


def bytes_shadow(rays, nodes, prims):    # same with a 1-byte result
    return 65 * rays + 64 * nodes + 80 * prims


def bytes_gather(q, depth, cand, sel):   # B_gather = 48 + 64*D_leaf + 24*C + 48*min(32,C) + 24 per query
    return 72 * q + 64 * depth + 24 * cand + 48 * sel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--split", default="samples", choices=["samples", "tiles"], help="how N > 1 GPUs share the frame")
    ap.add_argument("--spp", type=int, default=0, help="samples per step (default: the config's step_spp)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--photons", type=int, default=-1)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from gi_raytracer_b200 import build, host
    from gi_raytracer_b200 import dist as gd
    from gi_raytracer_b200.abi import GiStats, render_params
    from gi_raytracer_b200.capi import Context

    cfg = dict(CONFIGS[args.config])
    if args.photons >= 0:
        cfg["photons"] = args.photons
    W, H, MAX_DEPTH = cfg["w"], cfg["h"], cfg["depth"]
    SPP = args.spp or cfg["step_spp"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ctx = Context(local)          # raises without a B200: no CPU fallback
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    if world > 1:
        gd.init_comm(ctx, rank, world, dev)    # gi_comm_init: the C ABI's own NCCL communicator (torch carries the 128-byte id)

    def barrier_sync():
        ctx.synchronize()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    scene = host.load_scene(scene_file(cfg))
    if scene.n_prims == 0:
        raise RuntimeError("scene assets missing: run `make -C oracle assets` / `python scenes/make_standins.py`")
    ctx.upload_scene(scene)
    K, Wm = args.steps, max(args.warmup, 3)

    # -- photon phase (once, like the reference's cached map): rank 0 traces + builds, ONE slab broadcast (gi_photon_map_bcast)
    t0 = time.time()
    photon_stats = None
    if rank == 0:
        if cfg["photons"]:
            n_ph, photon_stats = ctx.photon_trace(cfg["photons"], 5, seed=1)
        else:
            ctx.photon_upload(np.zeros((0, 9)))
        ctx.photon_map_build(None)
    ctx.synchronize()
    tb0 = time.time()
    if world > 1:
        ctx.photon_map_bcast(0)
        ctx.synchronize()
    bcast_s = time.time() - tb0
    slab_bytes = ctx.photon_map_slab()[1]
    barrier_sync()
    photon_wall = time.time() - t0
    pm_info = ctx.photon_map_info()

    tiles = args.split == "tiles"
    P_tiles = render_params(W, H, SPP, max_depth=MAX_DEPTH, seed=1)               # the one frame every N renders under the tile split
    s0, s1 = gd.sample_ranges(SPP, world)[rank]
    P_samples = render_params(W, H, SPP * world, max_depth=MAX_DEPTH, seed=1)     # sample split: the frame grows with N (weak scaling)
    my_rows = ctx.rows_of_part(H, BLOCK_ROWS, world, rank) if H > rank * BLOCK_ROWS else 0
    # device buffers: two full-frame accumulators (the collective of frame i overlaps nothing it must not: same stream), local rows, images
    accums = [torch.zeros((H * W, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    rows_acc = torch.zeros((max(my_rows, 1) * W, 3), dtype=torch.float64, device=dev)
    rows_rgb = torch.zeros((max(my_rows, 1) * W, 3), dtype=torch.uint8, device=dev)
    frame_rgb = torch.zeros((H * W, 3), dtype=torch.uint8, device=dev)
    step_no = [0]
    no_reduce = bool(os.environ.get("GI_BENCH_NO_REDUCE"))   # A/B knob: the frames without the collective

    def step_samples():
        """sample split: own sample range of every pixel, then the fp64 partial sums are added onto rank 0 (same stream)"""
        accum = accums[step_no[0] & 1]
        step_no[0] += 1
        st = ctx.render_tile_dev(P_samples, 0, 0, W, H, s0, s1, accum.data_ptr())
        if world > 1 and not no_reduce:
            ctx.framebuffer_reduce(accum.data_ptr(), H * W * 3, 0)
        return st

    def step_tiles():
        """tile split: own interleaved row blocks in one wavefront, resolved locally, 8-bit rows gathered on rank 0 (same stream)"""
        st = ctx.render_rows_dev(P_tiles, BLOCK_ROWS, world, rank, 0, SPP, rows_acc.data_ptr())
        ctx.resolve_dev(my_rows * W, rows_acc.data_ptr(), SPP, rows_rgb.data_ptr())
        if world > 1 and not no_reduce:
            ctx.framebuffer_gather(rows_rgb.data_ptr(), W * 3, H, BLOCK_ROWS, frame_rgb.data_ptr(), 0)
        return st

    def timed_region(step, k):
        barrier_sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        sts = [step() for _ in range(k)]
        ev1.record(stream)
        barrier_sync()
        return ev0.elapsed_time(ev1), sts

    step_device = step_tiles if tiles else step_samples
    # clocks / throttle reasons are sampled from the first warm-up step to the end of the timed region (same load throughout);
    # nvidia-smi needs a moment to start, so warm-up continues (beyond W steps, at most 3 s) until it has delivered a sample
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_w = time.time()
    done = 0
    while True:
        step_device()
        done += 1
        ctx.synchronize()
        ready = torch.tensor([1 if (rank != 0 or clocks.lines or clocks.proc is None or time.time() - t_w > 3.0) else 0], device=dev)
        if world > 1:
            dist.broadcast(ready, src=0)
        if done >= Wm and int(ready.item()):
            break
    ms_total, stats = timed_region(step_device, K)
    clk = clocks.stop() if rank == 0 else None
    rays_local = sum(int(s.closest_rays) + int(s.shadow_rays) for s in stats)
    launches_local = sum(int(s.kernel_launches) for s in stats)

    # -- the other split at this N (N > 1, default line only): strong-scaling numbers of the same frame under the tile split
    other = None
    if world > 1 and not tiles:
        for _ in range(2):
            step_tiles()
        ms_o, st_o = timed_region(step_tiles, K)
        other = (ms_o, sum(int(s.closest_rays) + int(s.shadow_rays) for s in st_o))

    # -- per-family kernel durations for the roofline: the same frame on ONE stream, so that every kernel has the GPU to itself while
    #    its events are taken (inside the overlapped frame above the families' event times cover each other)
    ctx.configure("overlap_threshold", 0)
    ctx.synchronize()
    for _ in range(2):
        serial = ctx.render_tile_dev(P_samples, 0, 0, W, H, s0, s1, accums[0].data_ptr())
    ctx.synchronize()
    fam_launches = {k: ctx.kernel_ms(k)[1] for k in ("bounce", "direct", "gather")}
    ctx.configure("overlap_threshold", 1 << 20)

    # -- isolated gather: queries = primary hits (s = 0) of the frame, resident in HBM
    gather_ms, nq, gwork = 0.0, 0, [0, 0, 0, 0]
    if cfg["photons"]:
        o, d, _ = ctx.camera_rays(W, H, 0, 0, W, H, 0, 1)
        prim, hit, nrm, _ = ctx.trace_closest(o, d)
        m = prim != 0xFFFFFFFF
        nn = nrm[m].copy()
        flip = (nn * d[m]).sum(axis=1) > 0
        nn[flip] *= -1.0
        refl = d[m] - nn * (nn * d[m]).sum(axis=1)[:, None] * 2.0
        q_pos = torch.from_numpy(np.ascontiguousarray(hit[m])).to(dev)
        q_dir = torch.from_numpy(np.ascontiguousarray(refl)).to(dev)
        q_rgb = torch.empty_like(q_pos)
        nq = q_pos.shape[0]
        for _ in range(3):
            ctx.gather_dev(nq, q_pos.data_ptr(), q_dir.data_ptr(), q_rgb.data_ptr(), 32)
        barrier_sync()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        GREP = 10
        for _ in range(GREP):
            ctx.gather_dev(nq, q_pos.data_ptr(), q_dir.data_ptr(), q_rgb.data_ptr(), 32)
        g1.record(stream)
        barrier_sync()
        gather_ms = g0.elapsed_time(g1) / GREP
        gwork = ctx.last_work("gather")

    # -- e2e through the host-pointer C ABI (pinned host buffers).  Every step: the scene description in (gi_scene_upload on every rank),
    #    the frame, and ONE resolved image + its fp64 sums out on rank 0.  N > 1: the framebuffer collective sits between render and read-back.
    scene_bytes = sum(getattr(scene, f).nbytes for f in ("node_box", "node_child", "node_mask", "node_prim_off", "node_prim_cnt", "leaf_prims", "prim_type", "prim_geom",
                                                        "prim_nrm", "prim_uv", "prim_fnorm", "prim_mat", "mats", "tex", "tex_pixels", "lights"))
    rgb8 = torch.empty((H * W, 3), dtype=torch.uint8).pin_memory()
    acc_host = torch.empty((H * W, 3), dtype=torch.float64).pin_memory()
    e2e_d2h = H * W * 3 + H * W * 24

    def step_e2e():
        ctx.upload_scene(scene)
        if world == 1:   # what RayTracer::run calls per frame: rendered and resolved on the device, image + sums to the host
            st = GiStats()
            ctx._ck(ctx.L.gi_render_image(ctx.h, C.byref(P_samples), 0, 0, W, H, s0, s1, rgb8.numpy().ctypes.data, acc_host.numpy().ctypes.data, C.byref(st)))
            return st
        # (the read-backs run on torch's own stream after the context's stream has drained: a torch copy enqueued on the context's
        #  external stream would tie the tensors' lifetime to a stream that gi_destroy takes away)
        if tiles:
            st = step_tiles()
            ctx.synchronize()
            if rank == 0:
                rgb8.copy_(frame_rgb)
        else:
            st = step_samples()
            if rank == 0:
                accum = accums[(step_no[0] - 1) & 1]
                ctx.resolve_dev(H * W, accum.data_ptr(), SPP * world, frame_rgb.data_ptr())
            ctx.synchronize()
            if rank == 0:
                rgb8.copy_(frame_rgb)
                acc_host.copy_(accum)
        return st

    step_e2e()
    barrier_sync()
    te0 = time.perf_counter()
    e2e_stats = [step_e2e() for _ in range(K)]
    barrier_sync()
    e2e_s = time.perf_counter() - te0
    e2e_rays_local = sum(int(s.closest_rays) + int(s.shadow_rays) for s in e2e_stats)

    # -- max over ranks / sums over ranks
    red = torch.tensor([ms_total, gather_ms, e2e_s, other[0] if other else 0.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([rays_local, nq, e2e_rays_local, launches_local, other[1] if other else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, gather_ms, e2e_s, other_ms = [float(v) for v in red.cpu()]
    rays_all, nq_all, e2e_rays_all, launches_all, other_rays = [float(v) for v in tot.cpu()]

    if rank == 0:
        peak, peak_kind = load_peaks()
        last = stats[-1]
        g = lambda f: int(getattr(last, f))   # noqa: E731
        # the wavefront kernels' share of the work = totals minus what the tail kernel (one warp per path, deep bounces) did
        fam = {
            "bounce": (float(serial.trace_ms), bytes_closest(g("closest_rays") - g("tail_closest_rays"), g("closest_node_tests") - g("tail_closest_node_tests"),
                                                           g("closest_prim_tests") - g("tail_closest_prim_tests"))),
            "direct": (float(serial.shadow_ms), bytes_shadow(g("shadow_rays") - g("tail_shadow_rays"), g("shadow_node_tests") - g("tail_shadow_node_tests"),
                                                           g("shadow_prim_tests") - g("tail_shadow_prim_tests"))),
            "gather": (float(serial.gather_ms), bytes_gather(g("gathers") - g("tail_gathers"), g("gather_leaf_depth") - g("tail_gather_leaf_depth"),
                                                           g("gather_candidates") - g("tail_gather_candidates"), g("gather_selected") - g("tail_gather_selected"))),
        }
        if tiles and world > 1:   # the serial frame above is a FULL frame; under the tile split a rank renders 1 / N of it
            fam = {k: (v[0], v[1]) for k, v in fam.items()}
        tail_bytes = (bytes_closest(g("tail_closest_rays"), g("tail_closest_node_tests"), g("tail_closest_prim_tests"))
                      + bytes_shadow(g("tail_shadow_rays"), g("tail_shadow_node_tests"), g("tail_shadow_prim_tests"))
                      + bytes_gather(g("tail_gathers"), g("tail_gather_leaf_depth"), g("tail_gather_candidates"), g("tail_gather_selected")))
        sg = lambda f: int(getattr(serial, f))   # noqa: E731
        # the families' bytes belong to the frame their times were taken in (the serial full-sample-range frame of this rank)
        fam_bytes = {
            "bounce": bytes_closest(sg("closest_rays") - sg("tail_closest_rays"), sg("closest_node_tests") - sg("tail_closest_node_tests"), sg("closest_prim_tests") - sg("tail_closest_prim_tests")),
            "direct": bytes_shadow(sg("shadow_rays") - sg("tail_shadow_rays"), sg("shadow_node_tests") - sg("tail_shadow_node_tests"), sg("shadow_prim_tests") - sg("tail_shadow_prim_tests")),
            "gather": bytes_gather(sg("gathers") - sg("tail_gathers"), sg("gather_leaf_depth") - sg("tail_gather_leaf_depth"), sg("gather_candidates") - sg("tail_gather_candidates"),
                                   sg("gather_selected") - sg("tail_gather_selected")),
        }
        fam = {k: (fam[k][0], fam_bytes[k]) for k in fam}
        frame_bytes = sum(bytes_ for _, bytes_ in fam.values()) + tail_bytes
        dom = max(fam, key=lambda k: fam[k][0])
        dom_ms, dom_bytes = fam[dom]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = (tj.get(args.config) or {}).get(dom) if isinstance(tj.get(args.config), dict) else (tj.get(dom) if args.config == "C2" else None)
            except Exception:
                traffic = None
        gather_bytes = bytes_gather(gwork[0], gwork[1], gwork[2], gwork[3])
        famd = {k: {"ms_per_step": v[0], "launches_per_step": fam_launches[k], "algorithmic_bytes": v[1], "algorithmic_GBps": (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0),
                    "frac": (v[1] / (v[0] * 1e-3) / 1e9 / peak if v[0] > 0 else 0.0)} for k, v in fam.items()}
        roofline = {"bound": "hbm", "kernel": {"bounce": "k_bounce (closest hit + shade)", "direct": "k_direct (shadow any-hit)", "gather": "k_gather_* (locate + sorted + heavy)"}[dom],
                    "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes_per_step": dom_bytes, "kernel_ms_per_step": dom_ms, "launches_per_step": fam_launches[dom],
                    "families": famd,
                    "counts": "node / primitive tests are the device's own tallies = the canonical ordered traversal of SURVEY 8d with the walk's pruning rules R1-R3 (what the kernels execute; tests/test_gpu_parity.py checks them against the CPU count of the same rules) - the reference's unpruned walk tests more, so these bytes are the smaller, executed figure",
                    "timing": "families: CUDA events around each kernel family in one extra frame rendered on ONE stream right after the timed region (inside the timed, overlapped frames the families run beside each other and their event windows cover one another); frame: that single-stream frame; frame_timed: the timed region itself",
                    "families_overlapped_ms_per_step": {"bounce": float(last.trace_ms), "direct": float(last.shadow_ms), "gather": float(last.gather_ms)},
                    "frame_serial_ms": float(serial.total_ms),
                    "tail": {"ms_per_step": float(last.shade_ms), "algorithmic_GBps": tail_bytes / (float(last.shade_ms) * 1e-3) / 1e9 if last.shade_ms > 0 else 0.0,
                             "rays": g("tail_closest_rays") + g("tail_shadow_rays"), "gathers": g("tail_gathers")},
                    "bin_ms_per_step": float(last.bin_ms),
                    "frame": {"algorithmic_bytes": frame_bytes, "ms": float(serial.total_ms), "algorithmic_GBps": frame_bytes / (float(serial.total_ms) * 1e-3) / 1e9,
                              "frac": frame_bytes / (float(serial.total_ms) * 1e-3) / 1e9 / peak},
                    # the same bytes over the TIMED frame (three streams, families beside each other): what the schedule adds over the single-stream frame
                    "frame_timed": {"ms": ms_total / K, "algorithmic_GBps": frame_bytes / (ms_total / K * 1e-3) / 1e9, "frac": frame_bytes / (ms_total / K * 1e-3) / 1e9 / peak},
                    "gather_isolated": ({"ms": gather_ms, "queries": nq, "algorithmic_GBps": gather_bytes / (gather_ms * 1e-3) / 1e9, "frac": gather_bytes / (gather_ms * 1e-3) / 1e9 / peak}
                                        if nq else None)}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                res = run_reference_sample(args.config, cfg, 2)
                best = max(res["steps"], key=lambda s: s["mrays"])
                gq = [s["gather_mq"] for s in res["steps"] if s["gather_mq"] is not None]
                cpu = {"value": best["mrays"], "unit": "Mrays/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                       "gather_mqueries_s": max(gq) if gq else None, "photon_phase_s": res["photon_s"], "wall_s": res["wall_s"]}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: " + str(e)[:160]}
        frame_spp = SPP if tiles else SPP * world
        line = {
            "metric": "Mrays/s (all bounces)", "value": rays_all / (ms_total * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "reference scene assets as mounted (synthetic stand-ins for C4/C5, see config.scene); random-free inputs, counter PRNG seed 1",
            "config": {"workload": workload_name(args.config, cfg, SPP) + (" per GPU (sample-index split)" if not tiles else " (frame split into interleaved 16-row blocks)"),
                       "scene": f"scenes/{cfg['scene']}/{cfg['scene']}.scn ({cfg['note']})", "width": W, "height": H, "spp_per_step": SPP, "frame_spp": frame_spp,
                       "photons_stored": pm_info["n_kept"], "photon_map_nodes": pm_info["n_nodes"], "l2": "working set per step (path state ~2.9 GB) exceeds the 126 MB L2",
                       "streams": "three streams per context: generate / bounce kernels / binning / the tail kernel on the main stream (highest priority), every depth's shadow rays on side stream 0, every gather run on side stream 1, hit lists in a ring of eight (gi_ctx::sched_mode 1); roofline.families are timed in one extra single-stream frame",
                       "parallelism": (f"tile-split x{world} (gi_render_rows + gi_framebuffer_gather of 8-bit rows)" if tiles else f"sample-split x{world} (gi_framebuffer_reduce of fp64 sums)"),
                       "collective": ("none (N = 1)" if world == 1 else ("disabled (GI_BENCH_NO_REDUCE)" if no_reduce else "inside the timed region, on the render stream")),
                       "photon_phase_s": photon_wall, "photon_slab_bytes": slab_bytes, "photon_bcast_s": bcast_s if world > 1 else 0.0,
                       "rays_per_step": rays_all / K, "closest_rays_per_step": int(last.closest_rays), "shadow_rays_per_step": int(last.shadow_rays), "gathers_per_step": int(last.gathers)},
            "gather": ({"metric": "photon-gather Mqueries/s", "value": nq_all / (gather_ms * 1e-3) / 1e6, "unit": "Mqueries/s", "queries": nq_all,
                        "candidates_per_query": gwork[2] / max(gwork[0], 1)} if nq else None),
            "photons": {"tries": int(photon_stats.photon_tries) if photon_stats else None, "traces": int(photon_stats.closest_rays) if photon_stats else None,
                        "trace_ms": float(photon_stats.total_ms) if photon_stats else None},
            "e2e": {"value": e2e_rays_all / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes) * world,
                    "d2h_bytes_per_step": int(e2e_d2h if (world == 1 or not tiles) else H * W * 3),
                    "ms_per_step": 1e3 * e2e_s / K,
                    "calls": ("gi_scene_upload + gi_render_image (host pointers: scene arrays in, 8-bit image + fp64 sums out)" if world == 1 else
                              ("gi_scene_upload + gi_render_rows + gi_resolve + gi_framebuffer_gather -> one 8-bit host image on rank 0" if tiles else
                               "gi_scene_upload + gi_render_tile + gi_framebuffer_reduce + gi_resolve -> one 8-bit host image + fp64 sums on rank 0"))},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clk,
        }
        if other:
            line["tile_split"] = {"scaling": "strong", "metric": "Mrays/s (all bounces)", "value": other_rays / (other_ms * 1e-3) / 1e6, "ms_per_step": other_ms / K,
                                  "frame": f"{W}x{H}, {SPP} spp (the N = 1 frame), interleaved {BLOCK_ROWS}-row blocks, gather of the 8-bit rows inside the timed region"}
        emit(line)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        ctx.comm_destroy()
        dist.destroy_process_group()
    ctx.close()


def emit(obj):
    """The ONE JSON line of the run goes to the real stdout; everything else printed while the run lasted went to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1
if __name__ == "__main__":
    # libraries print to fd 1 from C (NCCL's version banner, the reference's own chatter): keep stdout for the JSON line alone
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    main()
