#!/usr/bin/env python
"""bench.py — headline benchmark: Mrays/s on house.toml 1920x1080, 8 bounces.

A "step" is one full render of the frame: `--spp` samples per pixel (default 256, the
configuration BASELINE.json quotes) of assets/scenes/house.toml at 1920x1080 with
max_bounces 8 under the synthetic 2048x1024 HDRI, on ONE GPU. With N GPUs every rank
renders `--spp` samples of the whole frame over a DISJOINT sample-index range
(rank r: [r*spp, (r+1)*spp)), then the accumulation buffers are sum-reduced onto rank 0
over NCCL — the sample split of SURVEY §8(e); per-GPU work is fixed, so scaling is "weak".

value      = rays traced by all ranks / device time (CUDA events on the launch stream,
             max over ranks), inputs resident in HBM. Rays are counted in-kernel:
             closest-hit cast_ray calls + NEE shadow rays (BASELINE.md counting rule).
e2e        = same metric through the public API with HOST buffers inside the timed
             region: scene + camera upload (host -> device), render, reduce, and the
             device -> host read of the W*H*4-float accumulation buffer.
roofline   = the dominant kernel, wave_walk_kernel (BVH walk): algorithmic bytes of the
             traversal (32 B per box test + 48 B per triangle test, counted
             exactly by the reference-order kernel on the same workload, untimed) per
             launch / the kernel's average launch duration, measured with CUDA events
             around every launch of the timed region (rt_set_profiling), against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json. The scene itself is
             L1/L2 resident; the kernel is issue-bound, which `roofline_issue` (whole
             step, algorithmic FP32 operations against 148 SM x 128 lanes x 2 x f_SM)
             and `roofline_step` (all algorithmic bytes of the step) put beside it.
cpu_baseline / --impl reference = the C++ oracle (a restatement of shader.wgsl; the Rust
             + wgpu reference cannot be built or run in this image) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WIDTH, HEIGHT, BOUNCES = 1920, 1080, 8
SCENE = os.path.join(ROOT, "assets", "scenes", "house.toml")
WORKLOAD = "house.toml 1920x1080 256spp 8 bounces, synthetic 2048x1024 HDRI (BASELINE.json configs[1])"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
SM_COUNT, FP32_LANES = 148, 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=256, help="samples per pixel per GPU per step")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--cpu-seconds", type=float, default=12.0,
                    help="target CPU time of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def algorithmic_bytes(c, pixels):
    """bytes any traversal in the reference's order must touch."""
    shaded = c["shaded_bounces"]
    return (32 * c["nodes_tested"] + 48 * c["tris_tested"] + 16 * c["spheres_tested"] +
            64 * c["planes_tested"] + 48 * c["tri_hits"] + 16 * c["hits"] +
            shaded * (2 * 16) + c["env_lookups"] * (4 * 16) + 32 * pixels)


def traversal_bytes(c):
    """The walk kernel's share of algorithmic_bytes: box tests + triangle tests."""
    return 32 * c["nodes_tested"] + 48 * c["tris_tested"]


def kernel_algorithmic_bytes(c, pixels):
    """algorithmic_bytes split over the kernels that do the work (the terms add up to it)."""
    shaded = c["shaded_bounces"]
    return {
        "wave_walk_kernel": traversal_bytes(c),
        # sphere / plane tests: the analytic pre-pass (the finish kernel runs it for camera rays)
        "wave_prepass_kernel": 16 * c["spheres_tested"] + 64 * c["planes_tested"],
        # per shaded hit: normals, material, alias entries and texels of the next-event sample
        "wave_shade_kernel": 48 * c["tri_hits"] + 16 * c["hits"] + shaded * (2 * 16 + 4 * 16),
        # sky lookups of escaped rays, accumulation
        "wave_finish_kernel": (c["env_lookups"] - shaded) * (4 * 16) + 32 * pixels,
    }


def algorithmic_flops(c):
    """FP32 operations, with S_bounce frozen at 250."""
    return (18 * c["nodes_tested"] + 45 * c["tris_tested"] + 30 * c["spheres_tested"] +
            40 * c["planes_tested"] + 250 * c["shaded_bounces"])


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_oracle_rate(rt, orc, scene, env, width, height, spp, threads=0):
    """Times the oracle on `spp` samples of the workload; returns (Mrays/s, seconds, rays)."""
    ora = orc.Oracle(scene, env, width, height, max_bounces=BOUNCES)
    t0 = time.perf_counter()
    _, c = ora.render(0, spp, threads=threads)
    dt = time.perf_counter() - t0
    rays = c["rays"] + c["shadow_rays"]
    return rays / dt / 1e6, dt, rays


def run_reference(args):
    """The reference arm: the oracle (CPU restatement of shader.wgsl) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import _pkg
    rt = _pkg.load()
    import oracle as orc
    scene = rt.Scene.load_toml(SCENE)
    env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])
    cores = os.cpu_count() or 1
    # Bounded sample: size one step at ~args.cpu_seconds / steps of CPU work.
    rate1, dt1, _ = cpu_oracle_rate(rt, orc, scene, env, args.width, args.height, 1)
    budget = max(2.0, args.cpu_seconds)
    spp = max(1, min(16, int(budget / max(dt1, 1e-3))))
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(rt, orc, scene, env, args.width, args.height, 1)
    total_rays, total_dt = 0, 0.0
    for _ in range(args.steps):
        _, dt, rays = cpu_oracle_rate(rt, orc, scene, env, args.width, args.height, spp)
        total_rays += rays
        total_dt += dt
    value = total_rays / total_dt / 1e6
    sample = f"{spp} of 256 spp of the full 1920x1080 frame per step (samples 0..{spp - 1})"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample,
                   "note": "CPU restatement of shader.wgsl (oracle port); the Rust/wgpu reference cannot be built here"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import _pkg
    rt = _pkg.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the backend has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner (and any NCCL_DEBUG lines) to stdout while the
        # communicator is created; stdout carries exactly one JSON line, so it points at stderr
        # until that is done.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    n_gpus = world

    width, height, spp = args.width, args.height, args.spp
    pixels = width * height
    scene = rt.Scene.load_toml(SCENE)
    env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])
    state = rt.State(scene, env, width, height, device=local_rank, max_bounces=BOUNCES)

    # The accumulation buffer is a torch tensor so the NCCL reduce can run on it, and the
    # backend launches on torch's current stream so torch events bracket its kernels.
    accum = torch.zeros((height, width, 4), dtype=torch.float32, device="cuda")
    state.bind_accum(accum.data_ptr())
    stream = torch.cuda.current_stream()
    state.set_stream(stream.cuda_stream)
    host_accum = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    sample_begin, _ = rt.samplesplit.weak_range(spp, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        accum.zero_()
        state.render_range(sample_begin, spp)
        if world > 1:
            rt.samplesplit.reduce_accumulation(accum, dst=0)

    def e2e_step():
        # Host buffers -> device (scene arrays, camera), render, reduce, device -> host.
        state.upload_scene(scene)
        state.update()
        accum.zero_()
        state.render_range(sample_begin, spp)
        if world > 1:
            rt.samplesplit.reduce_accumulation(accum, dst=0)
        host_accum.copy_(accum, non_blocking=True)
        stream.synchronize()

    # ---- exact work counters of one step (untimed, detailed counting on) ----
    state.set_detailed_counters(True)
    state.reset_counters()
    accum.zero_()
    state.render_range(sample_begin, spp)
    counters = state.counters()
    state.set_detailed_counters(False)
    rays_per_step = counters["rays"] + counters["shadow_rays"]

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()

    # ---- timed: device-resident ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches_before = state.launch_count()
    step_ms, kernel_ms = [], []
    for _ in range(args.steps):
        flush.fill_(1)  # evict L2 between timed iterations
        barrier()
        es, e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        es.record()
        accum.zero_()
        e0.record()
        state.render_range(sample_begin, spp)
        e1.record()
        if world > 1:
            rt.samplesplit.reduce_accumulation(accum, dst=0)
        e2.record()
        barrier()
        step_ms.append(es.elapsed_time(e2))
        kernel_ms.append(e0.elapsed_time(e1))
    launches = state.launch_count() - launches_before
    # One more step of the same region with CUDA events around every kernel launch
    # (rt_set_profiling): per-kernel durations for the roofline. Kept out of `value`'s
    # steps because creating ~10^4 events per step slows the host's launch loop.
    state.set_profiling(True)
    flush.fill_(1)
    barrier()
    device_step()
    barrier()
    profile = state.profile()
    state.set_profiling(False)
    clocks = sampler.stop() if sampler else None

    # ---- timed: end to end with host buffers ----
    for _ in range(2):
        e2e_step()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        barrier()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    total_ms = max_over_ranks(sum(step_ms))
    total_kernel_ms = max_over_ranks(sum(kernel_ms))
    total_e2e_ms = max_over_ranks(sum(e2e_ms))
    all_rays_per_step = sum_over_ranks(rays_per_step)
    all_paths_per_step = sum_over_ranks(counters["paths"])

    if rank == 0:
        value = all_rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
        e2e_value = all_rays_per_step * args.steps / (total_e2e_ms * 1e-3) / 1e6
        msamples = all_paths_per_step * args.steps / (total_ms * 1e-3) / 1e6

        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as fh:
                peaks = json.load(fh)
            hbm_peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        else:
            hbm_peak, peak_src, sm_max = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)", 1965.0
        render_s = total_kernel_ms / args.steps * 1e-3  # rank-0 view of one render (all its launches)
        alg_bytes = algorithmic_bytes(counters, pixels)
        alg_flops = algorithmic_flops(counters)
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
        fp32_peak = SM_COUNT * FP32_LANES * 2 * sm_mhz * 1e6 / 1e12
        achieved_tf = alg_flops / render_s / 1e12
        # per-kernel rooflines; `roofline` is the one of the dominant kernel (largest share of the
        # profiled device time)
        kernels = profile["kernels"]
        kernel_total_ms = sum(k["ms"] for k in kernels.values()) or 1.0
        traffic_by_kernel, traffic_note = {}, "no ncu capture committed"
        tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic_by_kernel, traffic_note = tj.get("kernels", {}), tj.get("note", "")
        per_kernel_bytes = kernel_algorithmic_bytes(counters, pixels)
        kernel_rooflines = {}
        for name, k in kernels.items():
            n_launch = max(1, k["launches"])
            launch_s = k["ms"] / n_launch * 1e-3
            bytes_per_launch = per_kernel_bytes.get(name, 0) * profile["renders"] / n_launch
            gbs = bytes_per_launch / launch_s / 1e9 if launch_s > 0 else 0.0
            kernel_rooflines[name] = {
                "share_of_device_time": k["ms"] / kernel_total_ms, "launch_ms": launch_s * 1e3,
                "launches_per_step": n_launch // max(1, profile["renders"]),
                "algorithmic_bytes_per_launch": int(bytes_per_launch), "achieved": gbs, "frac": gbs / hbm_peak,
                "traffic": traffic_by_kernel.get(name, {}).get("dram_bytes_per_launch")}
        dominant = max(kernel_rooflines, key=lambda n: kernel_rooflines[n]["share_of_device_time"])
        dom = kernel_rooflines[dominant]

        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if (width, height, spp) == (WIDTH, HEIGHT, 256)
                       else f"house.toml {width}x{height} {spp}spp 8 bounces (non-default size)",
                       "spp_per_gpu": spp, "parallelism": f"sample-split x{n_gpus} + NCCL sum-reduce",
                       "l2": "256 MiB flush write between timed steps",
                       "rays_per_step": int(all_rays_per_step),
                       "msamples_per_s": msamples},
            "e2e": {"value": e2e_value, "unit": "Mrays/s",
                    "h2d_bytes_per_step": scene_upload_bytes(scene) + 80,
                    "d2h_bytes_per_step": pixels * 16, "ms_per_step": total_e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": dom["achieved"], "peak": hbm_peak, "unit": "GB/s",
                         "frac": dom["frac"], "traffic": dom["traffic"],
                         "peak_source": peak_src, "kernel": dominant,
                         "kernel_share_of_device_time": dom["share_of_device_time"],
                         "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                         "launch_ms": dom["launch_ms"], "launches_per_step": dom["launches_per_step"],
                         "traffic_note": traffic_note,
                         "note": "the scene is L1/L2 resident, so the algorithmic bytes of the walk are served on "
                                 "chip (its fraction can exceed 1) and the kernels are bound by instruction issue "
                                 "and by the memory system's rate for scattered 32/64-byte pieces (roofline_issue, "
                                 "DESIGN.md section 5); the DRAM traffic is the wavefront's own state"},
            "roofline_kernels": kernel_rooflines,
            "roofline_step": {"bound": "hbm", "achieved": alg_bytes / render_s / 1e9, "peak": hbm_peak,
                              "unit": "GB/s", "frac": alg_bytes / render_s / 1e9 / hbm_peak,
                              "algorithmic_bytes_per_step": int(alg_bytes), "render_ms": render_s * 1e3},
            "roofline_issue": {"bound": "fp32-issue", "achieved": achieved_tf, "peak": fp32_peak,
                               "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
                               "sm_mhz": sm_mhz, "algorithmic_flops_per_step": int(alg_flops)},
            "kernel_profile": {"waves_per_step": profile["waves"] // max(1, profile["renders"]),
                               "gap_ms_per_step": profile["gap_ms"] / max(1, profile["renders"]),
                               "ms_per_step": {k: v["ms"] / max(1, profile["renders"]) for k, v in kernels.items()}},
            "counters_per_step_rank0": counters,
        }
        if not args.no_cpu_baseline:
            import oracle as orc
            cores = os.cpu_count() or 1
            _, dt1, _ = cpu_oracle_rate(rt, orc, scene, env, width, height, 1)
            n = max(1, min(96, int(args.cpu_seconds / max(dt1, 1e-3))))
            rate, dt, _ = cpu_oracle_rate(rt, orc, scene, env, width, height, n)
            out["cpu_baseline"] = {"value": rate, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                   "sample": f"{n} of {spp} spp of the full {width}x{height} frame "
                                             f"({dt:.1f} s of CPU work)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def scene_upload_bytes(scene):
    c = scene.counts()
    return (c["materials"] * 48 + c["spheres"] * 32 + c["planes"] * 96 + c["vertices"] * 16 +
            c["normals"] * 16 + c["triangles"] * 28 + c["primitives"] * 8 + c["bvh_nodes"] * 48)


if __name__ == "__main__":
    main()
