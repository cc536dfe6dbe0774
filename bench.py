#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 wavefront path tracer (BASELINE.json metric: Msamples/s and Mrays/s at
1/2/4/8 B200 vs the host-CPU reference; image RMSE vs CPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config bunny|cornell|glossy|large|cornell4k] [--quick]
    python bench.py --impl reference ...        # the reference's own CPU renderer, same metric / config

One "step" = one complete render of the workload: every pixel x spp samples through the hot path (generate, extend,
shade, connect, accumulate) + the film finalize.  With N > 1 (one process per GPU, torchrun) the scene is replicated,
the SAMPLE indices are partitioned, and the raw float32 films are summed onto rank 0 by ONE ncclReduce issued inside
the library (jpbrt_comm_init / jpbrt_read_film) -- torch.distributed only carries the rendezvous and the barriers.

The ONE JSON line rank 0 prints:
  value / ms_per_step : headline config (BASELINE configs[1], the bunny scene, 1024^2 x 50 spp PER GPU: weak scaling),
          scene resident in HBM, CUDA events on the library's stream, max over ranks.
  e2e   : the same through the calls a caller of FIntegrator::Render makes, with HOST buffers: every step re-uploads the
          flattened scene from pinned host memory, renders, reduces, finalizes, reads the film back (wall clock).
  strong: STRONG scaling -- the headline scene at 50 spp TOTAL and the 3840x2160 Cornell box at 4096 spp TOTAL
          (BASELINE configs[4]) split over the N ranks, with the film reduce timed on its own.
  configs (N = 1): every other BASELINE config (cornell, large at 64 spp, glossy at 64 spp) with value / e2e / roofline /
          cpu_baseline / image error, measured by the same code.
  roofline : the closest-hit traversal kernel (k_extend) against ceilings MEASURED in this run on this GPU by
          jet-pbrt_b200/build/peaks_l2 (scripts/peaks_l2.cu): ALGORITHMIC bytes -- the DISTINCT node (two boxes + two child
          references = 64 B) and primitive records each warp step needs (+ ray in, hit out) -- per second, over the gather
          bandwidth of 64-byte records at the scene's working-set size.  Where the tree is walked through the 32-byte quantised
          nodes (stats.node_bytes), `as_fetched` gives the bytes actually requested against the ceiling of that mix of 32- and
          64-byte records.  The per-lane SURVEY 8(d) figure, the HBM stream peak (MEASURED_PEAKS.json), the DRAM traffic and
          the issue-slot utilisation from this build's ncu capture (profiles/r02_ncu_constants.json) are quoted beside it.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref) on this host's cores, on a bounded sample of the same workload.
  reduce_check (N > 1): the NCCL-reduced film against all N x spp samples rendered by rank 0 alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONFIGS = {
    # name: (scene, scale, width, height, spp, BASELINE.json config it is)
    "bunny": ("bunny", 1.0, 1024, 1024, 50, "configs[1] bunny scene: 4 x 5,040-triangle mesh stand-in, matte/plastic/metal/glass, depth 5"),
    "cornell": ("cornell", 1.0, 1024, 1024, 50, "configs[0] Cornell box as in main.cc, depth 5"),
    "large": ("large", 1.0, 1024, 1024, 64, "configs[2] synthetic 5M-triangle scene, depth 8"),
    "glossy": ("glossy", 1.0, 1024, 1024, 64, "configs[3] glossy room, 16 area lights, depth 16"),
    "cornell4k": ("cornell", 1.0, 3840, 2160, 4096, "configs[4] Cornell box 3840x2160, 4096 spp"),
}
SEED = 1234


def workload_config(name, cfg, scene_d):
    """The workload, described identically by both arms (the driver compares the two `config` objects)."""
    scene_name, scale, w, h, spp, desc = cfg
    return {"workload": f"{name}: {desc}", "width": w, "height": h, "spp": spp, "max_depth": scene_d.max_depth,
            "n_primitives": scene_d.n_primitives, "n_lights": scene_d.n_lights, "seed": SEED}


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_peaks(device):
    """The gather / L2 / issue ceilings of THIS GPU (scripts/peaks_l2.cu), measured before the benchmark's own work."""
    exe = ROOT / "jet-pbrt_b200" / "build" / "peaks_l2"
    if not exe.exists():
        return None
    try:
        r = subprocess.run([str(exe), str(device)], capture_output=True, text=True, timeout=120)
        return json.loads(r.stdout) if r.returncode == 0 else None
    except Exception:
        return None


def gather_peak_at(peaks, working_set_bytes, key="divergent_gbs"):
    """Divergent 64-byte-record (key "divergent32_gbs": 32-byte-record) gather bandwidth at a working-set size:
    log-interpolated between the measured sizes."""
    import math

    pts = [(g["working_set_bytes"], g.get(key, g["divergent_gbs"])) for g in peaks["gather64"]]
    if working_set_bytes <= pts[0][0]:
        return pts[0][1]
    for (a, fa), (b, fb) in zip(pts, pts[1:]):
        if working_set_bytes <= b:
            t = (math.log(working_set_bytes) - math.log(a)) / (math.log(b) - math.log(a))
            return fa + t * (fb - fa)
    return pts[-1][1]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows, self.proc, self.device, self.seen = [], None, device, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])
            self.seen += 1

    def wait_first_sample(self, timeout_s: float):
        t0 = time.time()
        while self.proc and self.seen == 0 and time.time() - t0 < timeout_s:
            time.sleep(0.01)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# =====================================================================================================================
# reference arm: the reference's own CPU implementation of the path (FIntegrator::Render + parallel.cc thread pool)
# =====================================================================================================================
def load_scene_module():
    """jet-pbrt_b200/scene_desc.py + libjetpbrt_host.so: the scene DESCRIPTION only -- no CUDA library in this process."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("jpbrt_scene_desc", ROOT / "jet-pbrt_b200" / "scene_desc.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["jpbrt_scene_desc"] = mod
    spec.loader.exec_module(mod)
    mod.load_host_lib()
    return mod


def run_reference(args, name, cfg, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as ge

    sd = load_scene_module()
    orc = ge.load_oracle()
    scene_name, scale, w, h, spp, desc = cfg
    kind = "reference" if orc.have("ref") else "port"
    o = orc.Oracle("ref" if kind == "reference" else "port")
    threads = os.cpu_count() or 1
    sc = sd.HostScene.builtin(scene_name, w, h, scale)
    t0 = time.perf_counter()
    s = o.scene(sc)
    build_s = time.perf_counter() - t0
    # bounded sample: the step renders as many of the workload's spp as keep the whole K + W run near 3 minutes
    _, probe = s.render(1, threads)
    n_runs = args.steps + min(args.warmup, 1)
    step_spp = max(1, min(spp, int(min(180.0, 30.0 * n_runs) / n_runs / max(probe, 1e-3))))
    for _ in range(min(args.warmup, 1)):
        s.render(step_spp, threads)
    times = []
    for _ in range(args.steps):
        _, sec = s.render(step_spp, threads)
        times.append(sec)
    total = sum(times)
    value = w * h * step_spp * args.steps / total / 1e6
    line = {"impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, cfg, sc.d),
            "sample_spp_per_step": step_spp,
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": kind,
                             "sample": f"{w}x{h} x {step_spp} of the workload's {spp} spp per step (the metric is per sample), unmodified FIntegrator::Render with "
                                       f"numthreads={threads}; scene build {build_s:.2f}s excluded; no CUDA library loaded in this process"},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# =====================================================================================================================
# B200 arm
# =====================================================================================================================
class Bench:
    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist

        import __graft_entry__ as ge

        self.np, self.torch, self.dist, self.ge = np, torch, dist, ge
        self.args = args
        self.pkg = ge.load_package()
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available() or self.pkg.device_count() <= self.local:
            raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = None

    # ---- plumbing -------------------------------------------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def open(self, name, cfg_override=None):
        """Scene + context (+ the library's own NCCL communicator when N > 1)."""
        scene_name, scale, w, h, spp, desc = cfg_override or CONFIGS[name]
        sc = self.pkg.HostScene.builtin(scene_name, w, h, scale)
        t0 = time.perf_counter()
        ctx = self.pkg.Context(sc, device=self.local)
        upload_s = time.perf_counter() - t0
        if self.args.paths_in_flight:
            ctx.set_option("paths_in_flight", self.args.paths_in_flight)
        if self.world > 1:
            buf = self.torch.zeros(self.pkg.COMM_ID_BYTES, dtype=self.torch.uint8, device=self.dev)
            if self.rank == 0:
                uid = self.pkg.comm_unique_id()
                buf.copy_(self.torch.frombuffer(bytearray(uid), dtype=self.torch.uint8))
            self.dist.broadcast(buf, 0)
            ctx.comm_init(bytes(buf.cpu().numpy().tobytes()), self.rank, self.world)
        return sc, ctx, upload_s

    # ---- one configuration --------------------------------------------------------------------------------------------
    def measure(self, name, mode, steps, warmup, spp_override=0, full=True, sample_clocks=False, image_error=True, cpu=True, spp_warm=0):
        """mode "weak": every rank renders the config's spp (total = N x spp); "strong": the config's spp split over the ranks."""
        torch, np, pkg = self.torch, self.np, self.pkg
        cfg = list(CONFIGS[name])
        if spp_override:
            cfg[4] = spp_override
        scene_name, scale, w, h, spp, desc = cfg
        sc, ctx, upload_s = self.open(name, cfg)
        world, rank = self.world, self.rank
        if mode == "weak":
            begin, count, spp_total = rank * spp, spp, spp * world
        else:
            begin, count = pkg.sample_partition(spp, rank, world)
            spp_total = spp
        stream = torch.cuda.ExternalStream(ctx.stream(), device=self.local)
        nfloats = ctx.film_num_floats()
        host_film = torch.empty(nfloats, dtype=torch.float32, pin_memory=True) if rank == 0 else None
        host_ptr = host_film.data_ptr() if rank == 0 else None
        out = {}

        def step_resident(n=count, b=begin):
            ctx.clear_film()
            if n > 0:
                ctx.render_pass(b, n, SEED)
            ctx.reduce_film()                      # ncclReduce onto rank 0 on the library's stream (no-op at N = 1)
            if rank == 0:
                ctx.finalize_film_device(ctx.film_device_ptr(), spp_total)

        with torch.cuda.stream(stream):
            # ---- counting pass (outside every timed region): tests and DISTINCT fetches per ray on OUR BVH ----
            ctx.set_option("count_traversal", 1)
            ctx.clear_film()
            ctx.render_pass(begin, max(1, min(count, 4)), SEED)
            ctx.synchronize()
            cst = ctx.stats()
            ctx.set_option("count_traversal", 0)
            ctx.reset_stats()
            for _ in range(warmup):
                step_resident(spp_warm or count)
            self.barrier()

            # ---- e2e: host buffers in, host film out, every step (wall clock around the C-ABI calls) ----
            e2e = None
            if full:
                def step_e2e():
                    nbytes = ctx.reupload_scene()
                    ctx.clear_film()
                    if count > 0:
                        ctx.render_pass(begin, count, SEED)
                    ctx.read_film_root(spp_total, host_ptr)   # (reduce +) finalize + HBM -> pinned host, synchronises
                    return nbytes

                for _ in range(max(3, warmup)):
                    h2d = step_e2e()
                self.barrier()
                import gc
                gc.collect()
                gc.disable()
                t0 = time.perf_counter()
                each = []
                for _ in range(steps):
                    ts = time.perf_counter()
                    step_e2e()
                    each.append(1e3 * (time.perf_counter() - ts))
                self.barrier()
                e2e_s = self.max_over_ranks(time.perf_counter() - t0)
                gc.enable()
                e2e = {"value": w * h * spp_total * steps / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(nfloats * 4), "ms_per_step": 1e3 * e2e_s / steps, "ms_each_step": [round(x, 2) for x in each],
                       "what": "jpbrt_reupload_scene (pinned host -> HBM) + jpbrt_render_pass + jpbrt_read_film (ncclReduce at N > 1, finalize, HBM -> pinned host)"}
                out["film_mean"] = float(host_film.mean()) if rank == 0 else 0.0

            # ---- value: scene resident, CUDA events on the library's stream, per-stage events inside ----
            ctx.set_option("stage_timing", 1)  # per-launch CUDA events => the eager launch path (the graph path is what e2e ran)
            step_resident(spp_warm or count)
            self.barrier()
            clocks = ClockSampler(self.local)
            if sample_clocks and rank == 0 and not self.args.no_clocks:
                clocks.start()
                clocks.wait_first_sample(5.0)  # nvidia-smi's start-up stays OUT of the timed region
            self.barrier()
            step_resident(spp_warm or count)
            self.barrier()
            ctx.clear_film()
            ctx.reset_stats()
            clocks.rows.clear()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # hold the stream ~100 ms (untimed, before ev0) while the host queues the first steps: a stalled driver call
            # (seen on these boxes while nvidia-smi polls) cannot leave the GPU idle inside the timed region
            torch.cuda._sleep(int(0.1 * 1.9e9))
            ev0.record(stream)
            for _ in range(steps):
                step_resident()
            ev1.record(stream)
            self.barrier()
            ms_total = self.max_over_ranks(ev0.elapsed_time(ev1))
            clock_info = clocks.stop() if (sample_clocks and rank == 0 and not self.args.no_clocks) else None
            st = ctx.stats()
            ctx.set_option("stage_timing", 0)

            # ---- the film reduce on its own: every rank idle, then the collective alone (device time on rank 0) ----
            reduce_alone_ms = None
            if world > 1:
                times = []
                for _ in range(5):
                    ctx.clear_film()
                    ctx.render_pass(begin, 1, SEED)   # marks the film "not reduced"; its time is outside the events
                    self.barrier()
                    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    r0.record(stream)
                    ctx.reduce_film()
                    r1.record(stream)
                    self.barrier()
                    times.append(r0.elapsed_time(r1))
                reduce_alone_ms = self.max_over_ranks(min(times[1:]))

        ms_per_step = ms_total / steps
        ext_rays, sh_rays = st["extension_rays"], st["shadow_rays"]
        out.update({
            "value": w * h * spp_total / (ms_per_step * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
            "mode": mode, "spp_total": spp_total, "spp_this_rank": count,
            "mrays_per_s": (ext_rays + sh_rays) / steps * world / (ms_per_step * 1e-3) / 1e6,
            "rays_per_sample": (ext_rays + sh_rays) / steps / max(1, w * h * count),
            "stages_ms_per_step": {k[3:]: st[k] / steps for k in ("ms_generate", "ms_extend", "ms_shade", "ms_connect", "ms_finalize", "ms_reduce")},
            "reduce_alone_ms": reduce_alone_ms,
            "gpu_launches": int(st["kernel_launches"]), "paths_in_flight": int(st["paths_in_flight"]),
            "dropped": {k: int(st[k]) for k in ("invalid_contributions", "dropped_rays", "stack_overflows", "nee_dropped")},
            "upload_s": upload_s, "bvh_build_s": st["bvh_build_seconds"], "bvh_depth": int(st["bvh_depth"]), "scene_bytes": int(st["scene_bytes"]),
        })
        if e2e:
            out["e2e"] = e2e
        if clock_info:
            out["clocks"] = clock_info
        if rank == 0:
            out["roofline"] = self.roofline(name, st, cst, steps, sc.d.max_depth)
            out["config"] = workload_config(name, cfg, sc.d)
            if cpu and world == 1 and not self.args.no_cpu_baseline:
                out["cpu_baseline"] = self.cpu_baseline(cfg, sc)
                if image_error:
                    out["image_error_vs_cpu"] = self.image_error(cfg)
        ctx.close()
        return out

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------------
    def roofline(self, name, st, cst, steps, cst_depth=5):
        rays = max(cst["extension_rays"], 1)
        box, prim = cst["box_tests"] / rays, cst["prim_tests"] / rays
        nodef, primf = cst["node_fetches"] / rays, cst["prim_fetches"] / rays
        lane_bytes = 32.0 * box + 48.0 * prim + 48.0          # every lane's tests as if each were a fetch
        node_bytes = float(st.get("node_bytes") or 64)       # what a node step FETCHES: 64 (float nodes) or 32 (quantised nodes, mid-size trees)
        warp_bytes = 64.0 * nodef + 48.0 * primf + 48.0       # ALGORITHMIC: two 24-byte boxes + two child references per distinct node,
        #                                                        48 B per distinct primitive, ray in / hit out -- whatever the encoding
        fetched_bytes = node_bytes * nodef + 48.0 * primf + 48.0
        ext_ms = st["ms_extend"]
        n_rays = st["extension_rays"]
        achieved = n_rays * warp_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        lane_gbs = n_rays * lane_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        hbm, hbm_src = measured_hbm_peak()
        r = {"kernel": "k_extend (closest-hit BVH traversal)", "unit": "GB/s", "achieved": achieved,
             "kernel_ms_per_step": ext_ms / steps, "rays_per_step": int(n_rays / steps),
             "what": "ALGORITHMIC bytes per second of k_extend: the DISTINCT node records (two child boxes + references = 64 B) and 48-byte primitive records "
                     "each warp step needs + 48 B ray/hit per ray.  Mid-size trees are walked through 32-byte QUANTISED nodes (node_bytes): the same boxes in half "
                     "the bytes -- see as_fetched for the bytes actually requested",
             "node_bytes": int(node_bytes),
             "distinct_bytes_per_ray": warp_bytes, "node_fetches_per_ray": nodef, "prim_fetches_per_ray": primf,
             "per_lane_survey8d": {"bytes_per_ray": lane_bytes, "box_tests_per_ray": box, "prim_tests_per_ray": prim, "gbs": lane_gbs,
                                   "note": "32 B per box test + 48 B per primitive test + 48 B per ray, every lane counted (lanes on the same node share a fetch)"},
             "hbm_stream_peak": hbm, "hbm_peak_source": hbm_src}
        ws = int(st["scene_bytes"])
        cfgt = CONFIGS[name]
        n_bands = max(1, -(-(cfgt[2] * cfgt[3]) // (1 << 20)))
        band_pixels = -(-(cfgt[2] * cfgt[3]) // n_bands)
        spp_rank = max(1, int(st["samples"] / steps / (cfgt[2] * cfgt[3])))
        n_waves = n_bands * max(1, -(-spp_rank // max(1, int(st["paths_in_flight"]) // band_pixels)))
        launches = (int(cst_depth) + 1) * n_waves
        r["launches_per_step"] = launches
        if self.peaks:
            g = {x["working_set_bytes"]: x["divergent_gbs"] for x in self.peaks["gather64"]}
            l1_peak, l2_peak, dram_peak = g[min(g)], gather_peak_at(self.peaks, 32 << 20), g[max(g)]
            if ws <= (128 << 10):
                level, peak = "L1", l1_peak
            elif ws <= (96 << 20):
                level, peak = "L2", gather_peak_at(self.peaks, ws)
            else:  # the scene exceeds L2: no gather can beat the L2-resident rate; how much of it comes from DRAM is ncu's to say
                level, peak = "L2 (upper bound: the scene exceeds the 126 MB L2; a uniformly random gather over it reaches only %.0f GB/s)" % gather_peak_at(self.peaks, ws), l2_peak
            if node_bytes == 32 and "divergent32_gbs" in self.peaks["gather64"][0]:
                # what the quantised walk actually requests, against the ceiling of that mix: 32-byte records are ONE 256-bit load each;
                # the ceiling of the mix = the bytes over the time each part would take at its own ceiling
                peak32 = gather_peak_at(self.peaks, ws, "divergent32_gbs") if ws > (128 << 10) else self.peaks["gather64"][0]["divergent32_gbs"]
                t_min = node_bytes * nodef / peak32 + (48.0 * primf + 48.0) / peak
                fetched_gbs = n_rays * fetched_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
                r["as_fetched"] = {"bytes_per_ray": fetched_bytes, "gbs": fetched_gbs, "peak": fetched_bytes / t_min, "frac": fetched_gbs / (fetched_bytes / t_min),
                                   "gather_peak_32_byte_records_gbs": peak32, "gather_peak_64_byte_records_gbs": peak,
                                   "what": "the same with the node records at the 32 bytes they are fetched as, against the gather ceiling of that mix of 32- and 64-byte records"}
            r.update({"bound": f"issue + {level} gather of node and primitive records (scene working set {ws / 1e6:.1f} MB); ncu: IPC ~2.8 of 4 at ~15-21 of 32 lanes, DRAM < 10 %; with float nodes the L1/TEX pipe is the tighter of the two (77-84 % busy: 13 % fewer instructions per node step bought 0.3 %), with quantised nodes the issue rate (L1/TEX 46-56 %, IPC 2.9-3.0)",
                      "peak": peak, "frac": achieved / peak if peak else None,
                      "peak_source": "measured in this run by jet-pbrt_b200/build/peaks_l2 (scripts/peaks_l2.cu): divergent gather of 64-byte records (one per lane, two 256-bit loads) and of 32-byte records (one load) "
                                     "at the scene's working-set size; with quantised nodes the ceiling is that of the mix",
                      "gather_peaks_gbs": {"l1_resident": l1_peak, "l2_resident": l2_peak, "dram_random": dram_peak,
                                           "uniform_all_lanes_same_record": max(x["uniform_gbs"] for x in self.peaks["gather64"])},
                      "issue_peak_warp_ginst_s": self.peaks.get("ffma_warp_ginst_s"),
                      "frac_of_hbm_stream": achieved / hbm})
        else:
            r.update({"bound": "hbm", "peak": hbm, "frac": achieved / hbm, "peak_source": hbm_src + " (peaks_l2 micro-benchmark unavailable)"})
        consts = ROOT / "profiles" / "r02_ncu_constants.json"
        if consts.exists():
            try:
                c = json.loads(consts.read_text()).get("cornell" if name == "cornell4k" else name)
                if c:
                    ke = c["k_extend"]
                    r["traffic"] = ke["dram_bytes_per_ray"] * n_rays / steps / launches
                    r["traffic_unit"] = "bytes per launch: ncu dram__bytes_read.sum + dram__bytes_write.sum per ray of THIS build x rays per launch"
                    r["algorithmic_bytes_per_launch"] = n_rays * warp_bytes / steps / launches
                    issue = None
                    if self.peaks and ke.get("thread_inst_per_ray") and ext_ms > 0:
                        tps = ke["thread_inst_per_ray"] * n_rays / (ext_ms * 1e-3) / 1e9
                        issue = {"achieved_thread_ginst_s": tps, "peak_thread_ginst_s": 32.0 * self.peaks["ffma_warp_ginst_s"],
                                 "frac": tps / (32.0 * self.peaks["ffma_warp_ginst_s"]),
                                 "what": "thread instructions per ray (ncu, this build) x rays / k_extend time, over 32 lanes x the measured warp-instruction issue rate"}
                    r["ncu"] = {"source": c.get("source"), "k_extend": {k: ke.get(k) for k in ("dram_bytes_per_ray", "dram_frac", "l2_bytes_per_ray", "l2_gbs", "ipc", "lanes_per_inst",
                                                                                                "issue_lane_eff", "thread_inst_per_ray", "warp_inst_per_ray")},
                                "k_connect": {k: c.get("k_connect", {}).get(k) for k in ("dram_frac", "l2_gbs", "ipc", "lanes_per_inst", "issue_lane_eff")},
                                "issue": issue}
            except Exception as e:  # noqa: BLE001
                r["ncu_error"] = repr(e)
        r.setdefault("traffic", None)
        return r

    # ---- CPU reference on this host ---------------------------------------------------------------------------------------
    def cpu_baseline(self, cfg, sc):
        orc = self.ge.load_oracle()
        scene_name, scale, w, h, spp, desc = cfg
        kind = "reference" if orc.have("ref") else "port"
        o = orc.Oracle("ref" if kind == "reference" else "port")
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        s = o.scene(sc)
        build = time.perf_counter() - t0
        _, probe = s.render(1, threads)
        n = max(1, min(spp, int(self.args.cpu_seconds / max(probe, 1e-3))))
        sec = probe
        if n > 1:
            _, sec = s.render(n, threads)
        return {"value": w * h * n / sec / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
                "sample": f"{w}x{h} x {n} spp of the same scene, FIntegrator::Render numthreads={threads}, {sec:.1f}s (BVH build {build:.2f}s excluded)"}

    def image_error(self, cfg):
        """BASELINE.json metric, second half: image error of the GPU render against the reference CPU render at EQUAL spp, next to
        the CPU-vs-CPU figure between independently seeded reference renders (the Monte Carlo noise floor) -- K = 4 seeds on each
        side, all pairs, because a single pair scatters by several per cent.  Outside every timed region, at a reduced resolution so
        that the CPU side costs a few seconds; the comparisons at the configs' own size (1024^2 x 50 spp, K = 8) are in
        tests/test_gpu_full_size.py."""
        np, pkg = self.np, self.pkg
        orc = self.ge.load_oracle()
        scene_name, scale, w, h, spp, desc = cfg
        res, n, K = 256, 64, 4
        sc = pkg.HostScene.builtin(scene_name, res, max(1, res * h // w), scale if scene_name != "large" else 0.3)
        o = orc.Oracle("ref" if orc.have("ref") else "port").scene(sc)
        threads = os.cpu_count() or 1
        cpu = [o.render(n, threads, seed=1234 + 97 * k)[0] for k in range(K)]
        gpu = [pkg.render(sc, n, seed=5 + k, device=self.local)[0] for k in range(K)]
        rmse = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)))  # noqa: E731
        relmse = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))  # noqa: E731
        gc = [(g, c) for g in gpu for c in cpu]
        cc = [(cpu[i], cpu[j]) for i in range(K) for j in range(K) if i != j]
        out = {"what": f"{res}x{sc.d.camera.height} x {n} spp, clamped linear film values, K = {K} seeds per side, mean over all pairs",
               "rmse_gpu_vs_cpu": float(np.mean([rmse(a, b) for a, b in gc])), "rmse_cpu_vs_cpu": float(np.mean([rmse(a, b) for a, b in cc])),
               "relmse_gpu_vs_cpu": float(np.mean([relmse(a, b) for a, b in gc])), "relmse_cpu_vs_cpu": float(np.mean([relmse(a, b) for a, b in cc])),
               "mean_gpu": float(np.mean([g.mean() for g in gpu])), "mean_cpu": float(np.mean([c.mean() for c in cpu]))}
        out["relmse_ratio"] = out["relmse_gpu_vs_cpu"] / out["relmse_cpu_vs_cpu"]
        return out

    # ---- N > 1: the reduced film against one GPU rendering everything -------------------------------------------------------
    def reduce_check(self, name, spp):
        """SURVEY 8(e) validation ON THE B200s: ranks render their partitions, the library reduces; then rank 0 alone renders
        all N x spp sample indices.  The two raw films must agree to float summation order."""
        torch, np, pkg = self.torch, self.np, self.pkg
        sc, ctx, _ = self.open(name)
        world, rank = self.world, self.rank
        ctx.clear_film()
        ctx.render_pass(rank * spp, spp, SEED)
        ctx.reduce_film()
        ctx.synchronize()
        self.barrier()
        res = None
        if rank == 0:
            reduced = ctx.film_tensor().clone()
            ctx.clear_film()
            ctx.render_pass(0, spp * world, SEED)
            ctx.synchronize()
            alone = ctx.film_tensor()
            diff = (reduced - alone).abs()
            rel = diff / alone.abs().clamp_min(1e-3)
            res = {"what": f"{name} 1024^2: ncclReduce of {world} x {spp} spp partitions vs rank 0 rendering all {world * spp} spp alone, raw float32 sums",
                   "max_rel_err": float(rel.max()), "mean_rel_err": float(rel.mean()), "pixels_beyond_1e-5": int((rel > 1e-5).sum()),
                   "sum_reduced": float(reduced.double().sum()), "sum_alone": float(alone.double().sum()),
                   "ok": bool(rel.max() <= 1e-4 and abs(float(reduced.double().sum()) / float(alone.double().sum()) - 1) <= 1e-6)}
        self.barrier()
        ctx.close()
        return res


def run_b200(args, emit):
    args.warmup = max(args.warmup, 3)
    b = Bench(args)
    rank, world = b.rank, b.world
    if rank == 0:
        b.peaks = run_peaks(b.local)
    b.barrier()
    name = args.config
    head = b.measure(name, "weak", args.steps, args.warmup, spp_override=args.spp, full=True, sample_clocks=True)
    # ---- strong scaling: fixed total work split over the ranks ----
    strong = {}
    if not args.quick:
        s = b.measure(name, "strong", max(3, min(args.steps, 10)), 3, spp_override=args.spp, full=True, cpu=False)
        strong[name] = {k: s.get(k) for k in ("value", "unit", "ms_per_step", "spp_total", "spp_this_rank", "mrays_per_s", "stages_ms_per_step",
                                              "reduce_alone_ms", "e2e", "steps")}
        # BASELINE configs[4]: 3840 x 2160, 4096 spp in total, ONE timed render (18 s on one B200), warmed up with 3 short passes
        c5 = b.measure("cornell4k", "strong", 1, 3, spp_override=args.c5_spp, full=False, cpu=(world == 1), image_error=False, spp_warm=8)
        strong["cornell4k"] = {k: c5.get(k) for k in ("value", "unit", "ms_per_step", "spp_total", "spp_this_rank", "mrays_per_s", "rays_per_sample",
                                                      "stages_ms_per_step", "reduce_alone_ms", "roofline", "cpu_baseline", "config", "steps", "dropped")}
    # ---- the other BASELINE configs, one GPU ----
    configs = {}
    if world == 1 and not args.quick:
        for other in ("cornell", "large", "glossy"):
            if other == name:
                continue
            c = b.measure(other, "weak", 3, 3, full=True, image_error=(other != "large"))  # (a CPU image of the 5 M-triangle scene costs minutes)
            configs[other] = {k: c.get(k) for k in ("value", "unit", "ms_per_step", "e2e", "mrays_per_s", "rays_per_sample", "stages_ms_per_step", "roofline",
                                                    "cpu_baseline", "image_error_vs_cpu", "config", "upload_s", "bvh_build_s", "bvh_depth", "dropped", "steps")}
    check = b.reduce_check(name, 8) if world > 1 else None
    if rank == 0:
        cfg = CONFIGS[name]
        line = {
            "metric": "Msamples/s", "value": head["value"], "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": head["config"],
            "run": {"spp_per_gpu_per_step": head["spp_this_rank"], "spp_total": head["spp_total"],
                    "parallelism": f"sample-partition x{world}, scene replicated, ONE ncclReduce of the f32 film per step inside the library (jpbrt_comm_init / jpbrt_read_film)",
                    "paths_in_flight": head["paths_in_flight"],
                    "l2": f"no explicit flush: each step streams {cfg[2] * cfg[3] * head['spp_this_rank'] * 104 / 1e6:.0f} MB of wavefront state (> 126 MB L2); "
                          "the scene arrays are legitimately cache-resident across the step"},
            "mrays_per_s": head["mrays_per_s"], "rays_per_sample": head["rays_per_sample"],
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "stages_ms_per_step": head["stages_ms_per_step"], "reduce_alone_ms": head["reduce_alone_ms"],
            "clocks": head.get("clocks"), "dropped": head["dropped"],
            "upload_s": head["upload_s"], "bvh_build_s": head["bvh_build_s"], "film_mean": head.get("film_mean"),
            "peaks_measured": b.peaks,
        }
        for k in ("cpu_baseline", "image_error_vs_cpu"):
            if k in head:
                line[k] = head[k]
        if strong:
            line["strong"] = strong
        if configs:
            line["configs"] = configs
        if check:
            line["reduce_check"] = check
    if world > 1:
        b.dist.barrier()
        b.dist.destroy_process_group()
    if rank == 0:
        emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="bunny", choices=sorted(CONFIGS))
    ap.add_argument("--spp", type=int, default=0, help="override the config's samples per pixel")
    ap.add_argument("--c5-spp", type=int, default=0, help="override the 4096 spp of the 3840x2160 strong-scaling render")
    ap.add_argument("--quick", action="store_true", help="headline config only: no strong-scaling block, no other configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU time budget of each cpu_baseline sample")
    ap.add_argument("--paths-in-flight", type=int, default=0)
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi clocks during the timed region")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything else (NCCL banners, library chatter) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        return run_reference(args, args.config, CONFIGS[args.config], emit)
    return run_b200(args, emit)


if __name__ == "__main__":
    sys.exit(main())
