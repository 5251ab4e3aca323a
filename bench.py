#!/usr/bin/env python3
"""bench.py — the BASELINE metric on the BASELINE config.

Workload (N = 1): BASELINE config 2, `scenes/caustics` at 1024x1024, 8 fixed spp, MAX_DEPTH 64, 1 M caustic photons
(scenes/caustics/caustics.scn).  A step = one frame: the row loop of RayTracer::run (raytracer.h:93-160) over the whole
image with the photon map already built — the reference's own "second run()" timing — plus, reported beside it, the
isolated photon gather over the frame's primary-hit queries.

  metric  Mrays/s (all bounces) = (closest-hit + shadow traversals issued by the frame) / frame time
  value   whole-job throughput with everything resident in HBM (gi_render_tile_dev), CUDA events, max over ranks
  e2e     same metric through the host-pointer C ABI: gi_scene_upload + gi_render_image with HOST buffers
  gather  photon-gather Mqueries/s (gi_photon_gather_dev) over the primary-hit queries, same frame
  N > 1   weak scaling by sample index: rank r renders samples [8r, 8r+8) of every pixel; photon map built on rank 0
          and broadcast over NCCL (outside the timed region, like the reference's cached map); per-step NCCL reduce of
          the fp64 framebuffer sums (inside the timed region, double-buffered: it overlaps the next frame's rendering)

`--impl reference` times the reference's own OpenMP CPU path (oracle/_ref/gi_ref_fast, the unmodified reference
sources) on a bounded sample of the same workload, on the host cores of this box.
"""
import argparse
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

SCENE = os.path.join(ROOT, "scenes", "caustics", "caustics.scn")
W = H = 1024
SPP = 8
MAX_DEPTH = 64
PHOTONS = 1_000_000
CPU_SAMPLE_ROWS = (448, 512)   # the bounded CPU sample: 64 full rows of the 1024x1024 frame


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


# ---- the reference's CPU path (oracle/_ref) -------------------------------------------------------------------------------
def run_reference_sample(repeat, threads=None):
    """Run gi_ref_fast on the bounded sample; returns dict with per-repeat Mrays/s, gather Mq/s, core count."""
    exe = os.path.join(ROOT, "oracle", "_ref", "gi_ref_fast")
    if not os.path.exists(exe):
        raise FileNotFoundError(exe + " (build with `make -C oracle ref` where /root/reference is mounted)")
    threads = threads or os.cpu_count()
    out = tempfile.mkdtemp(prefix="gi_cpu_")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close")
    y0, y1 = CPU_SAMPLE_ROWS
    cmd = [exe, SCENE, out, "--w", str(W), "--h", str(H), "--y0", str(y0), "--y1", str(y1), "--s0", "0", "--s1", "1", "--repeat", str(repeat), "bench-frame", "time-gather"]
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
                      "gather_mq": meta["time_gather_queries"] / meta[f"time_gather_s_{r}"] / 1e6})
    return {"steps": steps, "cores": threads, "wall_s": wall, "photon_s": meta["photon_trace_s"] + meta["photon_build_s"], "photons": meta["photons_stored"],
            "sample": f"rows {y0}-{y1 - 1} of the {W}x{H} frame ({(y1 - y0) * W} pixels x {SPP} spp, MAX_DEPTH {MAX_DEPTH}, {PHOTONS} photons); gather: primary-hit queries of those rows"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, Wm = args.steps, args.warmup
    try:
        res = run_reference_sample(K + Wm)
    except Exception as e:  # the oracle always exists in a built tree; report honestly if the binary is missing
        emit({"impl": "reference", "unavailable": str(e).splitlines()[0][:200]})
        return
    timed = res["steps"][Wm:]
    tot_rays = sum(s["rays"] for s in timed)
    tot_s = sum(s["s"] for s in timed)
    value = tot_rays / tot_s / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s (all bounces)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": K, "warmup": Wm,
        "ms_per_step": 1e3 * tot_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 caustics 1024x1024, 8 spp, MAX_DEPTH 64, 1M photons (bounded sample per step)", "sample": res["sample"]},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                         "gather_mqueries_s": statistics.median(s["gather_mq"] for s in timed), "photon_phase_s": res["photon_s"]},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--photons", type=int, default=PHOTONS)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from gi_raytracer_b200 import build, host
    from gi_raytracer_b200 import dist as gd
    from gi_raytracer_b200.abi import render_params
    from gi_raytracer_b200.capi import Context

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

    def barrier_sync():
        ctx.synchronize()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    scene = host.load_scene(SCENE)
    if scene.n_prims == 0:
        raise RuntimeError("scene assets missing: run `make -C oracle assets` where /root/reference is mounted")
    ctx.upload_scene(scene)
    K, Wm = args.steps, max(args.warmup, 3)

    # -- photon phase (once, like the reference's cached map): rank 0 traces + builds, NCCL broadcast of the slab
    t0 = time.time()
    photon_stats = None
    if rank == 0:
        n_ph, photon_stats = ctx.photon_trace(args.photons, 5, seed=1)
        ctx.photon_map_build(None)
    slab_bytes = gd.share_photon_map(ctx, rank, world, 0)
    barrier_sync()
    photon_wall = time.time() - t0
    pm_info = ctx.photon_map_info()

    s0, s1 = gd.sample_ranges(SPP, world)[rank]
    P = render_params(W, H, SPP * world, max_depth=MAX_DEPTH, seed=1)
    # two framebuffers: the NCCL reduce of frame i (torch's stream) overlaps the rendering of frame i + 1 (the context's stream)
    accums = [torch.zeros((H * W, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    reduced = [None, None]   # event: the reduce that last read accums[b] is done
    step_no = [0]

    def step_device():
        b = step_no[0] & 1
        step_no[0] += 1
        accum = accums[b]
        if world > 1 and reduced[b] is not None:
            stream.wait_event(reduced[b])                         # the reduce of two frames ago still reads this buffer
        st = ctx.render_tile_dev(P, 0, 0, W, H, s0, s1, accum.data_ptr())
        if world > 1 and not os.environ.get("GI_BENCH_NO_REDUCE"):   # (debug knob: time the frames without the collective)
            torch.cuda.current_stream(dev).wait_stream(stream)
            gd.reduce_accum(accum, 0)
            reduced[b] = torch.cuda.Event()
            reduced[b].record(torch.cuda.current_stream(dev))
        return st

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
    barrier_sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    stats = []
    for _ in range(K):
        stats.append(step_device())
    if world > 1:
        stream.wait_stream(torch.cuda.current_stream(dev))   # the last reduces
    ev1.record(stream)
    barrier_sync()
    clk = clocks.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    rays_local = sum(int(s.closest_rays) + int(s.shadow_rays) for s in stats)
    launches_local = sum(int(s.kernel_launches) for s in stats)

    # -- per-family kernel durations for the roofline: the same frame on ONE stream, so that every kernel has the GPU to itself while
    #    its events are taken (inside the overlapped frame above the families' event times cover each other)
    ctx.configure("overlap_threshold", 0)
    ctx.synchronize()
    for _ in range(2):
        serial = ctx.render_tile_dev(P, 0, 0, W, H, s0, s1, accums[0].data_ptr())
    ctx.synchronize()
    ctx.configure("overlap_threshold", 1 << 20)

    # -- isolated gather: queries = primary hits (s = 0) of the frame, resident in HBM
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

    # -- e2e through the host-pointer C ABI (pinned host buffers): scene upload + frame + resolve, every step
    acc_host = torch.empty((H * W, 3), dtype=torch.float64).pin_memory().numpy()
    scene_bytes = sum(getattr(scene, f).nbytes for f in ("node_box", "node_child", "node_mask", "node_prim_off", "node_prim_cnt", "leaf_prims", "prim_type", "prim_geom",
                                                        "prim_nrm", "prim_uv", "prim_fnorm", "prim_mat", "mats", "tex", "tex_pixels", "lights"))
    import ctypes as C
    from gi_raytracer_b200.abi import GiStats
    rgb8 = torch.empty((H * W, 3), dtype=torch.uint8).pin_memory().numpy()

    def step_e2e():   # what RayTracer::run does per frame: scene description in, 8-bit image and fp64 sums out (pinned host buffers)
        ctx.upload_scene(scene)
        st = GiStats()
        ctx._ck(ctx.L.gi_render_image(ctx.h, C.byref(P), 0, 0, W, H, s0, s1, rgb8.ctypes.data, acc_host.ctypes.data, C.byref(st)))
        return st

    step_e2e()
    barrier_sync()
    te0 = time.perf_counter()
    e2e_stats = [step_e2e() for _ in range(K)]
    barrier_sync()
    e2e_s = time.perf_counter() - te0
    e2e_rays_local = sum(int(s.closest_rays) + int(s.shadow_rays) for s in e2e_stats)

    # -- max over ranks / sums over ranks
    red = torch.tensor([ms_total, gather_ms, e2e_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([rays_local, nq, e2e_rays_local, launches_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, gather_ms, e2e_s = [float(v) for v in red.cpu()]
    rays_all, nq_all, e2e_rays_all, launches_all = [float(v) for v in tot.cpu()]

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
        tail_bytes = (bytes_closest(g("tail_closest_rays"), g("tail_closest_node_tests"), g("tail_closest_prim_tests"))
                      + bytes_shadow(g("tail_shadow_rays"), g("tail_shadow_node_tests"), g("tail_shadow_prim_tests"))
                      + bytes_gather(g("tail_gathers"), g("tail_gather_leaf_depth"), g("tail_gather_candidates"), g("tail_gather_selected")))
        frame_bytes = sum(v[1] for v in fam.values()) + tail_bytes
        dom = max(fam, key=lambda k: fam[k][0])
        dom_ms, dom_bytes = fam[dom]
        n_launch = {"bounce": ctx.kernel_ms("bounce")[1], "direct": ctx.kernel_ms("direct")[1], "gather": ctx.kernel_ms("gather")[1]}
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)   # DRAM bytes per frame of the dominant family, from the committed ncu captures
            except Exception:
                traffic = None
        gather_bytes = bytes_gather(gwork[0], gwork[1], gwork[2], gwork[3])
        roofline = {"bound": "hbm", "kernel": {"bounce": "k_bounce (closest hit + shade)", "direct": "k_direct (shadow any-hit)", "gather": "k_gather"}[dom],
                    "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes_per_step": dom_bytes, "kernel_ms_per_step": dom_ms, "launches_per_step": n_launch[dom],
                    "families": {k: {"ms_per_step": v[0], "algorithmic_GBps": (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0), "frac": (v[1] / (v[0] * 1e-3) / 1e9 / peak if v[0] > 0 else 0.0)}
                                 for k, v in fam.items()},
                    "timing": "families: CUDA events around each kernel family in one extra frame rendered on ONE stream right after the timed region (inside the timed, overlapped frames the families run beside each other and their event windows cover one another); frame: the timed region itself",
                    "families_overlapped_ms_per_step": {"bounce": float(last.trace_ms), "direct": float(last.shadow_ms), "gather": float(last.gather_ms)},
                    "frame_serial_ms": float(serial.total_ms),
                    "tail": {"ms_per_step": float(last.shade_ms), "algorithmic_GBps": tail_bytes / (float(last.shade_ms) * 1e-3) / 1e9 if last.shade_ms > 0 else 0.0,
                             "rays": g("tail_closest_rays") + g("tail_shadow_rays"), "gathers": g("tail_gathers")},
                    "bin_ms_per_step": float(last.bin_ms),
                    "frame": {"algorithmic_bytes": frame_bytes, "ms": float(last.total_ms), "algorithmic_GBps": frame_bytes / (float(last.total_ms) * 1e-3) / 1e9,
                              "frac": frame_bytes / (float(last.total_ms) * 1e-3) / 1e9 / peak},
                    "gather_isolated": {"ms": gather_ms, "queries": nq, "algorithmic_GBps": gather_bytes / (gather_ms * 1e-3) / 1e9, "frac": gather_bytes / (gather_ms * 1e-3) / 1e9 / peak}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                res = run_reference_sample(2)
                best = max(res["steps"], key=lambda s: s["mrays"])
                cpu = {"value": best["mrays"], "unit": "Mrays/s", "cores": res["cores"], "kind": "reference", "sample": res["sample"],
                       "gather_mqueries_s": max(s["gather_mq"] for s in res["steps"]), "photon_phase_s": res["photon_s"], "wall_s": res["wall_s"]}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: " + str(e)[:160]}
        line = {
            "metric": "Mrays/s (all bounces)", "value": rays_all / (ms_total * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 scenes/caustics 1024x1024, 8 spp per GPU (sample-index split), MAX_DEPTH 64, 1M caustic photons, k=32 gather",
                       "scene": "scenes/caustics/caustics.scn (reference assets, dragon.obj not mounted)", "width": W, "height": H, "spp_per_gpu": SPP,
                       "photons_stored": pm_info["n_kept"], "photon_map_nodes": pm_info["n_nodes"], "l2": "working set per step (path state ~2.9 GB) exceeds the 126 MB L2",
                       "streams": "k_direct and the gather pipeline run on side streams (k_direct beside the gather; both behind the next depth's bounce kernel when a depth has < 2^20 hits); roofline.families are timed in one extra single-stream frame",
                       "parallelism": f"sample-split x{world}", "photon_phase_s": photon_wall, "photon_slab_bytes": slab_bytes,
                       "rays_per_step": rays_all / K, "closest_rays_per_step": int(last.closest_rays), "shadow_rays_per_step": int(last.shadow_rays), "gathers_per_step": int(last.gathers)},
            "gather": {"metric": "photon-gather Mqueries/s", "value": nq_all / (gather_ms * 1e-3) / 1e6, "unit": "Mqueries/s", "queries": nq_all,
                       "candidates_per_query": gwork[2] / max(gwork[0], 1)},
            "photons": {"tries": int(photon_stats.photon_tries) if photon_stats else None, "traces": int(photon_stats.closest_rays) if photon_stats else None,
                        "trace_ms": float(photon_stats.total_ms) if photon_stats else None},
            "e2e": {"value": e2e_rays_all / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(H * W * 24 + H * W * 3),
                    "ms_per_step": 1e3 * e2e_s / K, "calls": "gi_scene_upload + gi_render_image (host pointers: scene arrays in, 8-bit image + fp64 sums out)"},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clk,
        }
        emit(line)
    if world > 1:
        dist.barrier()
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
