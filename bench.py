#!/usr/bin/env python
"""bench.py — Mrays/s of the per-pixel path-tracing hot path on the BASELINE.json configs.

`--config` picks the workload (default B, the configuration the metric is quoted on):
  A  default.toml 640x360, 64 spp, 8 bounces           (the reference's own CPU-runnable case)
  B  house.toml 1920x1080, 256 spp, 8 bounces          (headline)
  C  suzanne.toml 3840x2160, primary rays only         (coherent rays; rt_first_hit)
  D  synthetic 99x99 field, ~10 M triangles, 1920x1080, 16 spp, 8 bounces   (scene > L2)
  E  house.toml 7680x4320, 1024 spp, 8 bounces         (multi-GPU sample split)

A "step" is one full render of the frame. With N GPUs (one process per GPU under torchrun)
the frame's samples are SPLIT: rank r renders `split_range(0, spp, r, N)` of the same frame
(shader.wgsl:1309-1312: a pixel-sample's seed depends only on (pixel, sample index)), then the
accumulation buffers are sum-reduced onto rank 0 over NCCL: total work is fixed, `"scaling":
"strong"` (32 spp per GPU at N = 8 for config B). `--scaling weak` keeps the
round-1 mode (every rank renders `spp` samples of its own range). Before the timed region an
N > 1 run renders a small frame split N ways, reduces it and compares it on rank 0 with that
rank's own render of the whole range (`reduce_check`); a mismatch fails the run.

value      = rays traced by all ranks (counted in-kernel by the timed kernels: closest-hit
             cast_ray calls + NEE shadow rays, BASELINE.md's rule) / device time (CUDA events
             on the launch stream, max over ranks), inputs resident in HBM.
e2e        = same metric through the public API with HOST buffers inside the timed region:
             scene + camera upload, render, reduce, device -> host read of the result.
clocks     = SM clock / power / throttle reasons sampled through in-process NVML every 500 ms from
             before the warm-up on; only samples inside the timed region are reported.
roofline   = configs A, B, C, E are bound by instruction issue: algorithmic FP32
             operations of the step (counted exactly by the reference-order kernel, untimed;
             18 / 45 / 30 / 40 per node / triangle / sphere / plane test + 250 per shaded
             bounce) per second, against 148 SMs x 128 lanes x f_SM — the build is
             --fmad=false, so one operation is one issued instruction. Config D (scene beyond
             L2) gets the HBM roofline: algorithmic bytes (32 B per box test + 48 B per
             triangle test + the per-hit / per-bounce terms) per second against the measured
             copy bandwidth. `roofline.measured` / `roofline.traffic` carry what ncu counted over
             every launch of the same step (profiles/measured.json, tools/ncu_step.py): warp and
             thread instructions, the share of the issue slots they filled, DRAM bytes.
cpu_baseline / --impl reference = the C++ oracle (a restatement of shader.wgsl; the Rust +
             wgpu reference cannot be built or run in this image) on all host cores, on a
             bounded sample of the same workload. kind = "port".
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

BOUNCES = 8
CONFIGS = {
    "A": dict(scene="default", width=640, height=360, spp=64, mode="path",
              workload="default.toml 640x360 64spp 8 bounces, synthetic 2048x1024 HDRI (BASELINE.json configs[0])"),
    "B": dict(scene="house", width=1920, height=1080, spp=256, mode="path",
              workload="house.toml 1920x1080 256spp 8 bounces, synthetic 2048x1024 HDRI (BASELINE.json configs[1])"),
    "C": dict(scene="suzanne", width=3840, height=2160, spp=1, mode="first_hit",
              workload="suzanne.toml 3840x2160 primary rays only, first-hit slot map (BASELINE.json configs[2])"),
    "D": dict(scene="synthetic:99", width=1920, height=1080, spp=16, mode="path",
              workload="synthetic field 99x99 instances = 9 997 020 triangles, 1920x1080 16spp 8 bounces "
                       "(BASELINE.json configs[3])"),
    "E": dict(scene="house", width=7680, height=4320, spp=1024, mode="path",
              workload="house.toml 7680x4320 1024spp 8 bounces, sample split + NCCL sum-reduce "
                       "(BASELINE.json configs[4])"),
}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
SM_COUNT, FP32_LANES = 148, 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--spp", type=int, default=0, help="override the config's samples per pixel")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0,
                    help="target CPU time of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def resolve_config(args):
    cfg = dict(CONFIGS[args.config])
    cfg["name"] = args.config
    changed = False
    for key in ("spp", "width", "height"):
        v = getattr(args, key)
        if v and v != cfg[key]:
            cfg[key] = v
            changed = True
    if changed:
        cfg["workload"] = (f"{cfg['scene']} {cfg['width']}x{cfg['height']} {cfg['spp']}spp 8 bounces "
                           f"(config {args.config} at a non-default size)")
    return cfg


def load_scene(rt, cfg, bvh_device=None, info=None):
    """bvh_device: CUDA device ordinal for the BVH of a generated scene (rt_bvh_build, byte-identical
    to the host builder); None builds it on the host like Scene::load_toml does."""
    if cfg["scene"].startswith("synthetic:"):
        grid = int(cfg["scene"].split(":")[1])
        if bvh_device is None:
            return rt.Scene.synthetic_field(os.path.join(ROOT, "assets"), grid)
        scene = rt.Scene.synthetic_field(os.path.join(ROOT, "assets"), grid, build_bvh=False)
        builder = rt.DeviceBvhBuilder(bvh_device)
        t0 = time.perf_counter()
        built = builder.build(scene)
        if info is not None:
            info.update({"where": f"cuda:{bvh_device} (rt_bvh_build)", "device_ms": built["device_ms"],
                         "with_upload_download_ms": (time.perf_counter() - t0) * 1e3,
                         "nodes": built["node_count"], "depth": built["depth"]})
        builder.close()
        return scene
    return rt.Scene.load_toml(os.path.join(ROOT, "assets", "scenes", f"{cfg['scene']}.toml"))


# ---- algorithmic work ------------------------------------------------------------
def algorithmic_bytes(c, pixels):
    """Bytes any traversal in the reference's order must touch."""
    shaded = c["shaded_bounces"]
    return (32 * c["nodes_tested"] + 48 * c["tris_tested"] + 16 * c["spheres_tested"] +
            64 * c["planes_tested"] + 48 * c["tri_hits"] + 16 * c["hits"] +
            shaded * (2 * 16) + c["env_lookups"] * (4 * 16) + 32 * pixels)


def algorithmic_ops(c):
    """FP32 operations, with S_bounce frozen at 250."""
    return (18 * c["nodes_tested"] + 45 * c["tris_tested"] + 30 * c["spheres_tested"] +
            40 * c["planes_tested"] + 250 * c["shaded_bounces"])


class ClockSampler:
    """Samples SM clock, power and the throttle reasons of one GPU while the timed region runs.

    In-process NVML (pynvml) from a thread: an `nvidia-smi -lms` child was measured to stall the
    renders of EVERY GPU of the box by 10-16 ms whenever one of its polls landed in a step (both
    ranks of an N = 2 run slowed down in the same steps; its start-up cost another 12 + 6 + 2 ms
    over the first three steps). The sampler therefore starts before the warm-up steps, polls every
    BENCH_SAMPLER_MS (default 500) ms, and only samples taken after mark() — the start of the timed
    region — are reported. Falls back to one nvidia-smi child if pynvml is unavailable."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))
    SMI_FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                  "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, uuid=None):
        self.interval = max(0.05, float(os.environ.get("BENCH_SAMPLER_MS", "500")) / 1e3)
        self.samples = []       # (t, sm_mhz, power_w, reason names)
        self.t_mark = None
        self.stop_flag = threading.Event()
        self.sm_max = None
        self.backend = None
        self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            if uuid:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
                except pynvml.NVMLError:
                    handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml, self.handle = pynvml, handle
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.backend = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
        except Exception:   # no pynvml / no NVML: one nvidia-smi child
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.SMI_FIELDS}",
                     "--format=csv,noheader,nounits", "-lms", str(int(self.interval * 1e3))],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.backend = "nvidia-smi"
                self.thread = threading.Thread(target=self._read_smi, daemon=True)
                self.thread.start()
            except OSError:
                self.backend = None

    def _poll_nvml(self):
        nv, h = self.nvml, self.handle
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                try:
                    power = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                except nv.NVMLError:
                    power = None
                names = [name for name, const in self.REASONS if mask & getattr(nv, const)]
                self.samples.append((time.perf_counter(), sm, power, names))
            except Exception:
                pass
            self.stop_flag.wait(self.interval)

    def _read_smi(self):
        names = [name for name, _ in self.REASONS]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm, mx, power = float(parts[0]), float(parts[1]), float(parts[2])
            except ValueError:
                continue
            self.sm_max = mx
            active = [n for n, v in zip(names, parts[3:7]) if v.lower().startswith("active")]
            self.samples.append((time.perf_counter(), sm, power, active))

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are not reported."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.backend:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"]}
        t_end = time.perf_counter()
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        else:
            self.thread.join(timeout=2)
        t0 = self.t_mark if self.t_mark is not None else 0.0
        inside = [s for s in self.samples if t0 <= s[0] <= t_end] or self.samples[-1:]
        sm = sorted(s[1] for s in inside)
        power = [s[2] for s in inside if s[2] is not None]
        reasons = sorted({n for s in inside for n in s[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                "power_w_max": max(power) if power else None, "samples": len(inside),
                "reasons": reasons, "source": f"{self.backend}, every {int(self.interval * 1e3)} ms"}


# ---- the CPU arm ------------------------------------------------------------------------------
def cpu_oracle_rate(orc, scene, env, cfg, spp, threads=0):
    """Times the oracle on `spp` samples of the workload; returns (Mrays/s, seconds, rays)."""
    ora = orc.Oracle(scene, env, cfg["width"], cfg["height"], max_bounces=BOUNCES)
    t0 = time.perf_counter()
    if cfg["mode"] == "first_hit":
        rays = 0
        for s in range(spp):
            ora.first_hit(s, threads=threads)
            rays += cfg["width"] * cfg["height"]
    else:
        _, c = ora.render(0, spp, threads=threads)
        rays = c["rays"] + c["shadow_rays"]
    dt = time.perf_counter() - t0
    return rays / dt / 1e6, dt, rays


def cpu_sample_text(cfg, n, dt):
    if cfg["mode"] == "first_hit":
        return f"{n} full {cfg['width']}x{cfg['height']} primary-ray frame(s) ({dt:.1f} s of CPU work)"
    return (f"{n} of {cfg['spp']} spp of the full {cfg['width']}x{cfg['height']} frame "
            f"(samples 0..{n - 1}, {dt:.1f} s of CPU work)")


def scene_buffers_check(scene, cfg):
    """The CPU arm takes its scene arrays from the product's host library; this pins them to the
    independent numpy restatement (oracle/scene_oracle.py) byte for byte before anything is timed."""
    if cfg["scene"].startswith("synthetic:"):
        return "not checked (10 M triangles are beyond the numpy restatement; tests/ cover small fields)"
    import numpy as np
    import scene_oracle
    path = os.path.join(ROOT, "assets", "scenes", f"{cfg['scene']}.toml")
    ref = scene_oracle.load_scene_toml(path)
    nodes, ordered = scene_oracle.build_bvh(ref)
    ok = scene.primitives_words().tolist() == [list(p) for p in ordered]
    words = scene.bvh_nodes_words()
    ok = ok and len(words) == len(nodes)
    if ok:
        for w, (mn, mx, idx, length, axis) in zip(words, nodes):
            f = w.view(np.float32)
            ok = ok and f[0:3].tobytes() == mn.tobytes() and f[4:7].tobytes() == mx.tobytes()
            ok = ok and (int(w[8]), int(w[9]), int(w[10])) == (idx, length, axis)
    ok = ok and scene.vertices().tobytes() == ref.vertices.tobytes()
    ok = ok and scene.normals().tobytes() == ref.normals.tobytes()
    ok = ok and scene.triangles_words().tolist() == [list(t) for t in ref.triangles]
    if not ok:
        raise SystemExit("scene buffers of librt_host.so differ from oracle/scene_oracle.py")
    return "BVH nodes, primitive order, vertices, normals, triangles byte-identical to oracle/scene_oracle.py"


def run_reference(args, cfg):
    """The reference arm: the oracle (CPU restatement of shader.wgsl) on all host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import _pkg
    rt = _pkg.load()
    import oracle as orc
    scene = load_scene(rt, cfg)
    check = scene_buffers_check(scene, cfg)
    env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])
    cores = os.cpu_count() or 1
    # Bounded sample: size one step at ~args.cpu_seconds of CPU work.
    _, dt1, _ = cpu_oracle_rate(orc, scene, env, cfg, 1)
    spp = max(1, min(cfg["spp"], 16, int(max(2.0, args.cpu_seconds) / max(dt1, 1e-3))))
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(orc, scene, env, cfg, 1)
    total_rays, total_dt = 0, 0.0
    for _ in range(args.steps):
        _, dt, rays = cpu_oracle_rate(orc, scene, env, cfg, spp)
        total_rays += rays
        total_dt += dt
    value = total_rays / total_dt / 1e6
    sample = cpu_sample_text(cfg, spp, total_dt / args.steps) + " per step"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": sample,
                   "note": "CPU restatement of shader.wgsl (oracle PORT, not the Rust/wgpu reference, which cannot "
                           "be built here); a rate metric, so the bounded sample stands for the whole frame",
                   "scene_check": check},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- the GPU arm ------------------------------------------------------------------------------
def main():
    args = parse_args()
    cfg = resolve_config(args)
    if args.impl == "reference":
        run_reference(args, cfg)
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

    width, height, spp = cfg["width"], cfg["height"], cfg["spp"]
    pixels = width * height
    first_hit = cfg["mode"] == "first_hit"
    t_build = time.perf_counter()
    bvh_info = {}
    scene = load_scene(rt, cfg, bvh_device=local_rank, info=bvh_info)
    build_s = time.perf_counter() - t_build
    t_env = time.perf_counter()
    env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])   # HDRI texels + alias table (environments.rs:97-186)
    env_s = time.perf_counter() - t_env
    t_state = time.perf_counter()
    state = rt.State(scene, env, width, height, device=local_rank, max_bounces=BOUNCES)
    state_s = time.perf_counter() - t_state   # State::new: context, scene re-layout + upload, environment upload
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    steps, warmup = args.steps, max(args.warmup, 3)

    def make_sampler():
        # rank 0 samples its own GPU; started before the warm-up steps so that its start-up cost
        # does not land in the timed region
        if rank != 0:
            return None
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            uuid = None
        return ClockSampler(local_rank, uuid)
    out = {"metric": "Mrays/s", "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic"}

    if first_hit:
        # ---- config C: primary rays only. Every rank renders the whole map (replicas). ----
        host_maps = (rt.pinned_array((height, width), np.uint32), rt.pinned_array((height, width), np.float32))
        sampler = make_sampler()
        for _ in range(warmup):
            state.first_hit(0, out=host_maps)
        barrier()
        if sampler:
            sampler.mark()
        launches_before = state.launch_count()
        kernel_ms, e2e_ms = [], []
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            t0 = time.perf_counter()
            state.first_hit(0, out=host_maps)       # kernel + 2 device -> host copies, synchronous
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
            kernel_ms.append(state.first_hit_kernel_ms())
            barrier()
        launches = state.launch_count() - launches_before
        clocks = sampler.stop() if sampler else None
        state.set_detailed_counters(True)
        state.reset_counters()
        state.first_hit(0, out=host_maps)
        counters = state.counters()
        state.set_detailed_counters(False)
        total_ms = max_over_ranks(sum(kernel_ms))
        total_e2e_ms = max_over_ranks(sum(e2e_ms))
        rays_per_step = sum_over_ranks(pixels)
        scaling = "weak"
        render_s = total_ms / steps * 1e-3
        extra_config = {"parallelism": f"replicas x{world} (the path is one frame of primary rays)",
                        "hit_fraction": float((host_maps[0] != 0xFFFFFFFF).mean())}
        e2e = {"value": rays_per_step * steps / (total_e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": 0, "d2h_bytes_per_step": pixels * 8, "ms_per_step": total_e2e_ms / steps,
               "note": "rt_first_hit into pinned host maps (rt_host_alloc): kernel + slot map + distance map copies"}
        profile = None
    else:
        # ---- path-traced configs ----
        if world > 1 and args.scaling == "strong":
            sample_begin, count = rt.samplesplit.split_range(0, spp, rank, world)
            scaling = "strong"
        else:
            sample_begin, count = rt.samplesplit.weak_range(spp, rank)
            scaling = "weak" if world > 1 else args.scaling   # one GPU: the two modes are the same run

        # The accumulation buffer is a torch tensor so the NCCL reduce can run on it, and the
        # backend launches on torch's current stream so torch events bracket its kernels.
        accum = torch.zeros((height, width, 4), dtype=torch.float32, device="cuda")
        state.bind_accum(accum.data_ptr())
        state.set_stream(stream.cuda_stream)
        host_accum = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()

        reduce_check = None
        if world > 1:
            reduce_check = run_reduce_check(rt, torch, dist, scene, env, local_rank, rank, world, stream)

        def render_and_reduce():
            accum.zero_()
            if count:
                state.render_range(sample_begin, count)
            if world > 1:
                rt.samplesplit.reduce_accumulation(accum, dst=0)

        # The scene bind group is part of the step's host inputs when it is small (the shipped scenes:
        # KBs). A scene of hundreds of MB (config D) is resident state, uploaded once like State::new
        # does (state.rs:394-458); its step uploads the camera uniform only, and the rate with the
        # upload inside the region is reported beside it.
        scene_bytes = scene_upload_bytes(scene)
        upload_scene_per_step = scene_bytes <= 64 << 20

        def e2e_step(with_scene=None):
            # Host buffers -> device (scene arrays, camera), render, reduce, device -> host.
            if upload_scene_per_step if with_scene is None else with_scene:
                state.upload_scene(scene)
            state.update()
            render_and_reduce()
            if rank == 0:      # the step's result is the reduced image, and it lives on rank 0
                host_accum.copy_(accum, non_blocking=True)
            stream.synchronize()

        sampler = make_sampler()
        for _ in range(warmup):
            render_and_reduce()
        barrier()

        # ---- timed: device-resident ----
        if sampler:
            sampler.mark()
        state.reset_counters()
        launches_before = state.launch_count()
        step_ms, kernel_ms = [], []
        for _ in range(steps):
            flush.fill_(1)  # evict L2 between timed iterations
            barrier()
            es, e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            es.record()
            accum.zero_()
            e0.record()
            if count:
                state.render_range(sample_begin, count)
            e1.record()
            if world > 1:
                rt.samplesplit.reduce_accumulation(accum, dst=0)
            e2.record()
            barrier()
            step_ms.append(es.elapsed_time(e2))
            kernel_ms.append(e0.elapsed_time(e1))
        launches = state.launch_count() - launches_before
        timed = state.counters()   # paths / rays / shadow rays of the timed kernels themselves
        if os.environ.get("BENCH_DEBUG"):
            print(f"[bench rank {rank}] step_ms {[round(v, 2) for v in step_ms]} render_ms "
                  f"{[round(v, 2) for v in kernel_ms]}", file=sys.stderr, flush=True)
        # One more step with per-kernel timing on (rt_set_profiling), kept out of `value`'s steps.
        state.set_profiling(True)
        flush.fill_(1)
        barrier()
        render_and_reduce()
        barrier()
        profile = state.profile()
        state.set_profiling(False)
        clocks = sampler.stop() if sampler else None

        # ---- timed: end to end with host buffers ----
        for _ in range(2):
            e2e_step()
        barrier()
        e2e_ms = []
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            barrier()
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        e2e_full_ms = None
        if not upload_scene_per_step:
            barrier()
            t0 = time.perf_counter()
            e2e_step(with_scene=True)
            barrier()
            e2e_full_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)

        # ---- exact algorithmic work (untimed): the reference-order kernel with detailed counting.
        # Bounded: on frames / sample counts beyond the headline's it runs a sample of this rank's range
        # and the figures are scaled by the ray count of the timed run.
        budget = 1920 * 1080 * 256
        detail_count = count if pixels * count <= budget else max(1, budget // pixels)
        counters, detail_note = dict(timed), "no samples on this rank"
        if count:
            state.set_detailed_counters(True)
            state.reset_counters()
            accum.zero_()
            state.render_range(sample_begin, detail_count)
            counters = state.counters()
            state.set_detailed_counters(False)
            per_step = {k: timed[k] // steps for k in ("paths", "rays", "shadow_rays")}
            if detail_count == count:
                if any(counters[k] != per_step[k] for k in per_step):
                    raise SystemExit(f"ray counters of the timed kernels {per_step} differ from the reference-order "
                                     f"kernel's {({k: counters[k] for k in per_step})}")
                detail_note = "exact: the reference-order kernel on the same samples; its ray counts equal the timed kernels'"
            else:
                scale = (per_step["rays"] + per_step["shadow_rays"]) / max(1, counters["rays"] + counters["shadow_rays"])
                counters = {k: int(v * scale) for k, v in counters.items()}
                detail_note = (f"reference-order kernel on {detail_count} of this rank's {count} samples, scaled by "
                               f"the timed kernels' ray count (x{scale:.3f})")
        total_ms = max_over_ranks(sum(step_ms))
        total_kernel_ms = max_over_ranks(sum(kernel_ms))
        total_e2e_ms = max_over_ranks(sum(e2e_ms))
        rays_per_step = sum_over_ranks((timed["rays"] + timed["shadow_rays"]) / steps)
        paths_per_step = sum_over_ranks(timed["paths"] / steps)
        render_s = total_kernel_ms / steps * 1e-3
        extra_config = {"parallelism": f"sample split x{world} + NCCL sum-reduce onto rank 0",
                        "samples_per_gpu": [rt.samplesplit.split_range(0, spp, r, world)[1] if scaling == "strong"
                                            else spp for r in range(world)],
                        "msamples_per_s": paths_per_step * steps / (total_ms * 1e-3) / 1e6,
                        "render_ms_rank_max": total_kernel_ms / steps,
                        "step_ms_rank0": [round(v, 3) for v in step_ms],
                        "reduce_ms": (total_ms - total_kernel_ms) / steps,
                        "algorithmic_work": detail_note}
        if reduce_check is not None:
            out["reduce_check"] = reduce_check
        e2e = {"value": rays_per_step * steps / (total_e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": (scene_bytes if upload_scene_per_step else 0) + 80,
               "d2h_bytes_per_step": pixels * 16, "ms_per_step": total_e2e_ms / steps}
        if world > 1:
            e2e["note"] = "every rank uploads scene and camera; the reduced image is read back on rank 0"

        if e2e_full_ms is not None:
            e2e["note"] = (f"the {scene_bytes / 1e6:.0f} MB scene is resident (uploaded once, as State::new does); a step "
                           f"uploads the camera uniform. With the scene upload and its re-layout inside the region: "
                           f"{e2e_full_ms:.0f} ms per step = {rays_per_step / (e2e_full_ms * 1e-3) / 1e6:.0f} Mrays/s")

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as fh:
                peaks = json.load(fh)
            hbm_peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        else:
            hbm_peak, peak_src, sm_max = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)", 1965.0
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
        # rank 0's share of the step, on rank 0's render time
        ops, abytes = algorithmic_ops(counters), algorithmic_bytes(counters, pixels)
        issue_peak = SM_COUNT * FP32_LANES * sm_mhz * 1e6 / 1e12   # one op per lane and clock: no FMA in this build
        issue = {"bound": "issue", "achieved": ops / render_s / 1e12, "peak": issue_peak, "unit": "Tops/s",
                 "frac": ops / render_s / 1e12 / issue_peak, "traffic": None,
                 "algorithmic_ops_per_step": int(ops), "render_ms": render_s * 1e3, "sm_mhz": sm_mhz,
                 "peak_source": "148 SMs x 128 FP32 lanes x median SM clock under load; the build is --fmad=false "
                                "(arithmetic contract, DESIGN.md section 2), so an operation is one issued "
                                "instruction and the FMA-doubled 74 TFLOP/s does not apply",
                 "kernel": "the whole render (one persistent launch, or the four wave kernels it is cut into)"}
        hbm = {"bound": "hbm", "achieved": abytes / render_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": abytes / render_s / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
               "algorithmic_bytes_per_step": int(abytes), "render_ms": render_s * 1e3}
        if cfg["name"] == "D":
            hbm["note"] = ("achieved counts ALGORITHMIC bytes (SURVEY §8d), most of which L1 / L2 serve; `traffic` is "
                           "the DRAM traffic ncu measured for the step. The deep walk is bound by divergence and "
                           "dependent-load latency (11.4 of 32 lanes), not by bandwidth (DESIGN.md section 5)")
        measured = load_measured_profile(cfg["name"])
        if measured:
            m = dict(measured.get("issue") or {})
            if m.get("warp_instructions_per_step") and cfg["spp"] == CONFIGS[cfg["name"]]["spp"] and world == 1:
                # ncu-counted instructions of the same step (profiles/measured.json): how busy the issue
                # ports were over the timed render, and how much of what they issued was algorithmic work
                warp_peak = SM_COUNT * 4 * sm_mhz * 1e6      # one warp instruction per scheduler and clock
                m["issue_slots_used_frac"] = m["warp_instructions_per_step"] / render_s / warp_peak
                m["algorithmic_share_of_thread_instructions"] = ops / m["thread_instructions_per_step"]
                m["active_threads_per_instruction"] = m["thread_instructions_per_step"] / m["warp_instructions_per_step"]
            issue["measured"] = m
            hbm["traffic"] = measured.get("dram_bytes_per_step")
            issue["traffic"] = measured.get("dram_bytes_per_step")
            hbm["traffic_source"] = issue["traffic_source"] = measured.get("source")
        out.update({
            "value": rays_per_step * steps / (total_ms * 1e-3) / 1e6,
            "ms_per_step": total_ms / steps, "scaling": scaling,
            "config": dict({"workload": cfg["workload"], "config": cfg["name"],
                            "l2": "256 MiB flush write between timed steps",
                            "rays_per_step": int(rays_per_step), "scene_build_s": build_s,
                            "environment_build_s": env_s, "state_new_s": state_s,
                            **({"bvh_build": bvh_info} if bvh_info else {})}, **extra_config),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": hbm if cfg["name"] == "D" else issue,
            "roofline_other": issue if cfg["name"] == "D" else hbm,
            "counters_per_step_rank0": counters,
        })
        if profile and profile["renders"]:
            out["kernel_profile"] = {
                "waves_per_step": profile["waves"] // profile["renders"],
                "launches_per_step": {k: v["launches"] // profile["renders"] for k, v in profile["kernels"].items()},
                "ms_per_step": {k: v["ms"] / profile["renders"] for k, v in profile["kernels"].items()},
                "gap_ms_per_step": profile["gap_ms"] / profile["renders"]}
        if not args.no_cpu_baseline:
            import oracle as orc
            cores = os.cpu_count() or 1
            _, dt1, _ = cpu_oracle_rate(orc, scene, env, cfg, 1)
            n = max(1, min(cfg["spp"], 96, int(args.cpu_seconds / max(dt1, 1e-3))))
            rate, dt, _ = cpu_oracle_rate(orc, scene, env, cfg, n)
            out["cpu_baseline"] = {"value": rate, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                   "sample": cpu_sample_text(cfg, n, dt)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reduce_check(rt, torch, dist, scene, env, local_rank, rank, world, stream):
    """A small frame split `world` ways and sum-reduced over NCCL must equal rank 0's own render of
    the whole sample range up to float re-association (SURVEY §8d parity gate 4)."""
    w, h, spp = 480, 270, 8 * world
    st = rt.State(scene, env, w, h, device=local_rank, max_bounces=BOUNCES)
    part = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    st.bind_accum(part.data_ptr())
    st.set_stream(stream.cuda_stream)
    begin, count = rt.samplesplit.split_range(0, spp, rank, world)
    st.render_range(begin, count)
    rt.samplesplit.reduce_accumulation(part, dst=0)
    stream.synchronize()
    result = None
    if rank == 0:
        whole = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        st.bind_accum(whole.data_ptr())
        st.render_range(0, spp)
        stream.synchronize()
        diff = (part[..., :3] - whole[..., :3]).abs().max().item()
        scale = whole[..., :3].abs().max().item()
        alpha_ok = bool((part[..., 3] == 1.0).all().item())
        tol = 1e-5 * spp * max(1.0, scale)
        result = {"frame": f"{w}x{h}", "spp": spp, "ranks": world, "max_abs_diff": diff, "tolerance": tol,
                  "max_value": scale, "alpha_is_one": alpha_ok, "ok": bool(diff <= tol and alpha_ok)}
    st.close()
    flag = torch.tensor([1 if (result is None or result["ok"]) else 0], device="cuda")
    dist.broadcast(flag, src=0)
    if int(flag.item()) != 1:
        if rank == 0:
            print(json.dumps({"error": "reduce_check failed", "reduce_check": result}))
        raise SystemExit(3)
    return result


def load_measured_profile(config_name):
    """ncu-measured figures of the committed capture for this config (profiles/measured.json),
    written by tools/ncu_summary.py from the .ncu-rep of the same command."""
    path = os.path.join(ROOT, "profiles", "measured.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        return json.load(fh).get(config_name)


def scene_upload_bytes(scene):
    c = scene.counts()
    return (c["materials"] * 48 + c["spheres"] * 32 + c["planes"] * 96 + c["vertices"] * 16 +
            c["normals"] * 16 + c["triangles"] * 28 + c["primitives"] * 8 + c["bvh_nodes"] * 48)


if __name__ == "__main__":
    main()
