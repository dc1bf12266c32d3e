#!/usr/bin/env python
"""Benchmark of the Level-S2fM render hot path (BASELINE.json metric: rendered rays/sec, forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--regime init|trained] [--impl ours|reference]

One "step" = one optimisation iteration's render work on one batch of synthetic rays: depth sampling ->
fused field kernel (hash grid + SDF MLP + normals + radiance) -> compositing -> loss (10^3 L1 rgb + 10^2 eikonal)
-> backward (compositing backward + fused field backward); for N > 1 the single all-reduce of the flat gradient
bucket is inside the step.  Workloads (``--workload``; BASELINE.json configs):

    c2 (default)  configs[1]: 4096 rays, error-bounded sampler 64 + 64, L=16, SDF 35-64-64-64-17, RadF 49-64-64-3, DTU bounds
    c1k           the same networks on the north star's 1024-ray batch
    uniform128    c2's networks, 128 uniform samples per ray
    c3            configs[2]: 8192 rays from 2 cameras, ETH3D bounds (+-5, inside: false), shipped networks, 128 uniform samples
    c4            configs[3]: 4096 rays, DTU, shipped networks + one sphere_tracing call per iteration
    c5            configs[4]: 16384 rays per ITERATION split over the N ranks (STRONG scaling), BlendedMVS bounds
    dual          shipped DTU configuration with Ablate_config.dual_field (scripts/train_DTU.sh:18)
    image         forward only: a whole 1200 x 1600 image in rand_rays slices (Camera.render_img_by_slices) -> rays/s
    grid          forward only: the 512^3 marching-cubes SDF volume (utils/util.py extract_mesh) -> points/s

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of every field.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GRID_BYTES_PER_EVAL = 1024      # L = 16 levels * 8 corners * F = 2 * 4 B  

C2_NETS = {"SDF.arch.layers": [None, 64, 64, 64, 16], "RadF.arch.layers": [None, 64, 64, 3]}
SHIPPED_NETS = {"SDF.arch.layers": [None, 64, 16], "RadF.arch.layers": [None, 64, 64, 3]}
UNIFORM = {"SDF.VolSDF.volsdf_sampling": False, "SDF.VolSDF.sample_intvs": 128}
EB = {"SDF.VolSDF.volsdf_sampling": True, "SDF.VolSDF.sample_intvs": 64, "SDF.VolSDF.final_sample_intvs": 64}

WORKLOADS = {
    "c2": dict(dataset="DTU", over={**C2_NETS, **EB}, rays=4096, cams=1, scaling="weak",
               desc="BASELINE configs[1]: 4096 rays, error-bounded sampler (64 coarse + 64 fine = 128 samples/ray), L=16 hash grid, "
                    "SDF MLP 35-64-64-64-17, RadF 49-64-64-3, fused fwd+bwd, DTU bounds"),
    "c1k": dict(dataset="DTU", over={**C2_NETS, **EB}, rays=1024, cams=1, scaling="weak",
                desc="north-star 1024-ray batch: error-bounded sampler (64 + 64 samples/ray), L=16 hash grid, SDF MLP 35-64-64-64-17, "
                     "RadF 49-64-64-3, fused fwd+bwd, DTU bounds"),
    "uniform128": dict(dataset="DTU", over={**C2_NETS, **UNIFORM}, rays=4096, cams=1, scaling="weak",
                       desc="4096 rays, 128 uniform samples/ray, L=16 hash grid, SDF MLP 35-64-64-64-17, RadF 49-64-64-3, fused fwd+bwd, DTU bounds"),
    "c3": dict(dataset="ETH3D", over={**SHIPPED_NETS, **UNIFORM}, rays=8192, cams=2, scaling="weak",
               desc="BASELINE configs[2] shape: 8192 rays/iter from 2 cameras (two-view init), 128 uniform samples/ray, ETH3D bounds +-5 "
                    "(inside: false, bias 2.5, scale_mlp 5), L=16 hash grid, shipped SDF MLP 35-64-17, RadF 49-64-64-3, eikonal + colour loss"),
    "c4": dict(dataset="DTU", over={**SHIPPED_NETS, **UNIFORM}, rays=4096, cams=1, scaling="weak", trace=True,
               desc="BASELINE configs[3] shape: 4096 rays, 128 uniform samples/ray, L=16 hash grid, shipped SDF MLP 35-64-17, RadF 49-64-64-3, "
                    "one sphere_tracing call on all rays per iteration (pipelines/Camera.py:506), fused fwd+bwd, DTU bounds"),
    "c5": dict(dataset="bmvs", over={**SHIPPED_NETS, **UNIFORM}, rays=16384, cams=1, scaling="strong",
               desc="BASELINE configs[4] shape: 16384 rays per ITERATION split over the ranks (strong scaling), 128 uniform samples/ray, "
                    "BlendedMVS bounds +-2 (bias 1, scale_mlp 3, white background), L=16 hash grid, shipped SDF MLP 35-64-17, RadF 49-64-64-3"),
    "dual": dict(dataset="DTU", over={**SHIPPED_NETS, **UNIFORM, "Ablate_config.dual_field": True}, rays=4096, cams=1, scaling="weak",
                 desc="shipped DTU configuration with dual_field (scripts/train_DTU.sh:18): 4096 rays, 128 uniform samples/ray, TWO L=16 hash "
                      "fields (SDF + RadF.Geo_enc, 35-64-17 each), RadF 65-64-64-3, fused fwd+bwd"),
}
FORWARD_ONLY = {
    "image": "forward only (SURVEY 8f row 4): one 1200 x 1600 image in rand_rays = 8192-ray slices as Camera.render_img_by_slices "
             "(pipelines/Camera.py:275-311), shipped DTU networks, 128 uniform samples/ray",
    "grid": "forward only (SURVEY 8f row 4): the 512^3 SDF volume of utils/util.py:392-430 (extract_mesh), shipped DTU SDF network, "
            "points generated on the device, values-only tensor-core kernel",
    "ba_sfm": "SURVEY 8f row 3: the point-only BA iteration of pipelines/BA.py:117-151 (surface projection, SDF at the projected points, "
              "se3 -> SE3 per observation, robust reprojection loss, L1 sdf + eikonal; forward + backward) on 4096 tracked points, "
              "shipped DTU SDF network, replayed as one CUDA graph",
}
METRIC = "rendered rays/sec (fwd+bwd, 4096-ray batch)"


def workload_opt(name: str, device: str):
    from levels2fm_b200.config import default_opt
    w = WORKLOADS[name]
    return default_opt(w["dataset"], device=device, **w["over"])


def bytes_per_ray(wl: dict, n_samples: int, k_rounds: float = 0.0, n_trace: float = 0.0) -> float:
    """SURVEY 8(d) / BASELINE.md section 3: fields * (S_eval * G + S_grad * G) + 64 B of ray I/O."""
    G = GRID_BYTES_PER_EVAL
    fields = 2 if wl["over"].get("Ablate_config.dual_field") else 1
    s_eval = n_samples
    if wl["over"].get("SDF.VolSDF.volsdf_sampling"):
        s_eval = 64 * (1 + k_rounds) + n_samples          # sampler passes (SDF field only) + final samples
        return (s_eval * G) + (fields - 1) * n_samples * G + fields * n_samples * G + 64
    return fields * (s_eval * G + n_samples * G) + 64 + n_trace * G


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a thread of this process
    (nvidia-smi -lms needs ~0.5 s to start, longer than a 20-step timed region); falls back to nvidia-smi when NVML is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.samples = []       # (sm_mhz, reasons bitmask)
        self.nvml = None
        self.stop_flag = False
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                                     int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))))
            except Exception:
                try:
                    self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                                         int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))))
                except Exception:
                    pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            n = self.nvml
            masks = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [s for s, _ in self.samples]
            reasons = sorted(k for k, m in masks.items() if any(r & m for _, r in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 5 ms poll during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


# ----------------------------------------------------------------------------- oracle arms (CPU reference / GPU eager)
def oracle_cfg(workload: str):
    """The oracle's scene description of a workload (same option tree as the product arm)."""
    from oracle import port
    return port.cfg_from_opt(workload_opt(workload, "cpu"))


def oracle_setup(workload: str, n_rays: int, seed: int = 0, device: str = "cpu", regime: str = "init"):
    """Oracle state with the synthetic scene of BASELINE.md section 3 (geometric sphere init, table U(-1e-4, 1e-4) /
    N(0, 0.05)), rays and gt colours for one step of ``n_rays`` rays."""
    from levels2fm_b200 import synthetic
    from oracle import port
    wl = WORKLOADS[workload]
    cfg = oracle_cfg(workload)
    half = float(cfg.bound_max[0])
    bias = {"DTU": 0.5, "ETH3D": 2.5, "bmvs": 1.0}[wl["dataset"]]
    sdf_sd, rad_sd = port.random_state(cfg, seed=0, table_std=1e-4 if regime == "init" else 0.05, generic_weights=False,
                                       sphere_bias=bias, hash_weight_std=0.0 if regime == "init" else 0.05)
    for sd in (sdf_sd, rad_sd):
        for k in sd:
            sd[k] = sd[k].to(device).requires_grad_(True)
    H, W = {"DTU": (1200, 1600), "ETH3D": (1033, 1551), "bmvs": (576, 768)}[wl["dataset"]]
    cams = wl["cams"]
    center, ray = synthetic.make_rays(cams, max(n_rays // cams, 1), half, H, W, seed=seed)
    gt = torch.rand(cams, max(n_rays // cams, 1), 3, generator=torch.Generator().manual_seed(seed + 1))
    return cfg, sdf_sd, rad_sd, center.to(device), ray.to(device), gt.to(device), bool(wl.get("trace"))


def oracle_step(cfg, sdf_sd, rad_sd, center, ray, gt, trace):
    from levels2fm_b200 import synthetic
    from oracle import port
    for sd in (sdf_sd, rad_sd):
        for v in sd.values():
            v.grad = None
    out = port.render_forward(center, ray, sdf_sd, rad_sd, cfg)
    loss = synthetic.render_loss(out, gt)
    if trace:
        st = port.sphere_tracing(center, ray, sdf_sd, cfg)
        loss = loss + 1e-2 * (st["d_pred"] - out["depth_mlp"][..., 0].detach()).abs().mean()
    loss.backward()
    return loss.detach()


def cpu_baseline(workload: str, budget_s: float = 12.0, n_rays: int = 128):
    """The oracle port (reference algorithm restated in eager PyTorch, CPU) on a bounded sample of the workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = oracle_setup(workload, n_rays)
    oracle_step(*st)                                   # warm-up
    t0, n = time.perf_counter(), 0
    while True:
        oracle_step(*st)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 20:
            break
    dt = time.perf_counter() - t0
    return {"value": n * n_rays / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{n} fwd+bwd iterations of {n_rays} rays of the same workload (oracle/port.py, torch CPU, {cores} threads)"}


def gpu_eager_baseline(workload: str, sdf, rad, center, ray, gt, ours_loss: float, steps: int = 5):
    """SURVEY 8(d) "reference timing beside it (2)": the same oracle (the reference's algorithm in eager PyTorch, PyTorch hash
    grid -- NOT tcnn, which cannot be installed here) on the same B200, stepping the FULL batch of the workload with the product
    arm's own weights, rays and gt colours -- so the two losses must agree (``loss_rel_diff``: the "matched loss" of the north
    star at step 0).  This is the stand-in for "the reference's single-GPU rays/sec" of the >= 10x target."""
    cfg = oracle_cfg(workload)
    sdf_sd = {k: v.detach().clone().requires_grad_(True) for k, v in sdf.state_dict().items()}
    rad_sd = {k: v.detach().clone().requires_grad_(True) for k, v in rad.state_dict().items()}
    st = (cfg, sdf_sd, rad_sd, center, ray, gt, bool(WORKLOADS[workload].get("trace")))
    for _ in range(2):
        loss = oracle_step(*st)
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = oracle_step(*st)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = statistics.median(a.elapsed_time(b) for a, b in evs)
    n_rays = center.shape[0] * center.shape[1]
    return {"value": n_rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "steps": steps, "rays_per_step": n_rays,
            "loss": float(loss), "ours_loss": ours_loss, "loss_rel_diff": abs(float(loss) - ours_loss) / abs(float(loss)),
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2),
            "what": "oracle/port.py on cuda: the reference's algorithm in eager PyTorch with a pure-PyTorch hash grid (not tcnn / vren), "
                    "same workload, same batch, same weights / rays / gt as the product arm, fwd + loss + bwd"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: tcnn / vren cannot be installed)
    on the host cores, every step a bounded 128-ray sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_rays = 128
    wl = WORKLOADS[args.workload]
    st = oracle_setup(args.workload, n_rays, regime=args.regime)
    for _ in range(max(args.warmup, 1)):
        oracle_step(*st)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle_step(*st)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = n_rays / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "rays_per_gpu": wl["rays"], "regime": args.regime,
                       "sample": f"each step = {n_rays} rays of that workload (CPU, bounded)"},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} timed fwd+bwd steps of {n_rays} rays (oracle/port.py: the reference's algorithm "
                                       "restated in eager PyTorch on the host cores; tcnn/vren are not installable here)"},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- forward-only workloads
def run_forward_only(args):
    from levels2fm_b200 import _C, ops, rays as rays_mod, synthetic
    from levels2fm_b200.config import default_opt
    from levels2fm_b200.models.RadF import RadF
    from levels2fm_b200.models.Renderer import Renderer
    from levels2fm_b200.models.SDF import SDF
    _C.get()
    dev = "cuda:0"
    torch.cuda.set_device(0)
    opt = default_opt("DTU", device=dev, **{**SHIPPED_NETS, **UNIFORM})
    torch.manual_seed(0)
    sdf, rad, ren = SDF(opt).to(dev), RadF(opt).to(dev), Renderer(opt)
    synthetic.init_fields(sdf, rad, args.regime)
    W = max(args.warmup, 3)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    clocks = ClockSampler(0)
    if args.workload == "ba_sfm":
        return run_ba_sfm(args, opt, sdf, dev, peak, clocks, W)
    if args.workload == "image":
        H, Wd = opt.data.image_size
        rot, pos = synthetic.look_at_cameras(1, 1.0, torch.Generator().manual_seed(0))
        pose = torch.cat([rot[0].T, (-rot[0].T @ pos[0])[:, None]], dim=1)[None].to(dev)          # world -> camera [R | t]
        f = 1.2 * Wd
        intr = torch.tensor([[f, 0.0, Wd / 2], [0.0, f, H / 2], [0.0, 0.0, 1.0]], device=dev)
        center, ray = rays_mod.get_center_and_ray(opt, pose, intr=intr[None])
        units, unit_name, per_unit = H * Wd, "rays/s", 128 * GRID_BYTES_PER_EVAL + 128 * 28

        def step():
            return ren.render_image(opt, center, ray, sdf, rad)["rgb"]
    else:
        N = 512
        units, unit_name, per_unit = N ** 3, "points/s", GRID_BYTES_PER_EVAL + 4

        def step():
            return sdf.infer_sdf_grid(N=N, volume_size=2.0)
    for _ in range(W):
        out = step()
    torch.cuda.synchronize()
    clocks.start()
    ops.KLOG.reset()
    evs = []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = step()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    clk = clocks.stop()
    launches = ops.KLOG.total()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    # e2e: the result leaves the device every step (the mesh exporter / image writer consumes it on the host)
    host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host.copy_(step(), non_blocking=True)
    torch.cuda.synchronize()
    e2e = units * args.steps / (time.perf_counter() - t0)
    achieved = units * per_unit / (ms * 1e-3) / 1e9
    line = {"metric": f"{'rendered rays' if args.workload == 'image' else 'SDF grid points'}/sec (forward only)", "value": units / (ms * 1e-3),
            "unit": unit_name, "n_gpus": 1, "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": FORWARD_ONLY[args.workload], "units_per_step": units, "regime": args.regime,
                       "l2": "inputs larger than L2 are not needed: the 48.8 MB table is meant to stay L2-resident across slices"},
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": unit_name, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(out.numel() * 4)},
            "roofline": {"bound": "hbm", "kernel": "field_forward", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "algorithmic_bytes_per_unit": per_unit,
                         "note": "whole step (all launches), algorithmic bytes = gather + per-sample outputs"}}
    print(json.dumps(line))


def run_ba_sfm(args, opt, sdf, dev, peak, clocks, W):
    from levels2fm_b200 import ba, ops, parallel
    from levels2fm_b200.graph import GraphedStep
    n = 4096
    g = torch.Generator().manual_seed(0)
    xyz = torch.nn.Parameter((torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * (0.5 + 0.02 * torch.randn(n, 1, generator=g))).to(dev))
    se3 = torch.nn.Parameter(torch.tensor([[0.05, -0.1, 0.02, 0.05, -0.03, 2.5], [-0.2, 0.3, 0.1, -0.1, 0.02, 2.6]], device=dev))
    pose_idx = (torch.arange(n) % 2).to(dev)
    intr = torch.tensor([[1920.0, 0.0, 800.0], [0.0, 1920.0, 600.0], [0.0, 0.0, 1.0]], device=dev)
    kp_h = (torch.rand(n, 2, generator=g) * torch.tensor([1600.0, 1200.0])).pin_memory()
    kp = kp_h.to(dev)
    bucket = parallel.GradBucket(list(sdf.parameters()) + [xyz, se3])
    thr = 2.0 / 10 / 100

    def iteration(kp_in):
        bucket.zero()
        t = ba.surface_ba_terms(sdf, xyz, se3, pose_idx, intr, kp_in, thr)
        loss = t["reproj_loss"] + 100.0 * t["sdf_surf"] + 100.0 * t["eikonal_loss"]        # 10 ** loss_weight.ba (options/LevelS2fM.yaml:113-118)
        loss.backward()
        return loss.detach()

    def timed(fn, steps):
        evs = []
        for _ in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn(kp)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / steps, out
    for _ in range(W):
        iteration(kp)
    torch.cuda.synchronize()
    ops.KLOG.reset()
    eager_ms, _ = timed(iteration, args.steps)
    launches_per_it = ops.KLOG.total() // args.steps
    step = GraphedStep(iteration, (kp,))
    for _ in range(W):
        step(kp)
    clocks.start()
    graph_ms, loss = timed(step, args.steps)
    clk = clocks.stop()
    host = torch.empty(1).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host.copy_(step(kp_h.to(dev, non_blocking=True)).reshape(1), non_blocking=True)
    torch.cuda.synchronize()
    e2e = n * args.steps / (time.perf_counter() - t0)
    per_pt = 2 * (2 * GRID_BYTES_PER_EVAL) + 64          # two field evaluations, each gathered and scattered once
    achieved = n * per_pt / (graph_ms * 1e-3) / 1e9
    line = {"metric": "BA sfm tracked points/sec (fwd+bwd, 4096 points per iteration)", "value": n / (graph_ms * 1e-3), "unit": "points/s",
            "n_gpus": 1, "steps": args.steps, "warmup": W, "ms_per_step": graph_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": FORWARD_ONLY["ba_sfm"], "points_per_iteration": n, "cuda_graph": True,
                       "eager_ms_per_iteration": eager_ms, "graph_speedup": eager_ms / graph_ms, "launches_per_iteration": launches_per_it,
                       "l2": "the iteration is launch / latency bound (4096 points): no L2 flush"},
            "clocks": clk, "gpu_launches": launches_per_it * args.steps, "loss": float(loss),
            "e2e": {"value": e2e, "unit": "points/s", "h2d_bytes_per_step": int(kp_h.numel() * 4), "d2h_bytes_per_step": 4},
            "roofline": {"bound": "hbm", "kernel": "whole iteration", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "algorithmic_bytes_per_unit": per_pt,
                         "note": "launch-bound by construction: 4096 points are 0.3 % of one C2 render batch"}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + sorted(FORWARD_ONLY))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regime", default="init", choices=["init", "trained"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--symmetric", choices=["auto", "on", "off"], default="auto",
                    help="N > 1: gradient bucket in symmetric memory + NVLS multimem all-reduce (in-switch reduction); auto = on from "
                         "8 ranks (measured, 48.9 MB bucket: 0.17 vs 0.27 ms at 8 GPUs, 0.27 vs 0.19 ms at 4, 0.43 vs 0.15 ms at 2)")
    ap.add_argument("--e2e-sync-readback", action="store_true", help="diagnostic: read the loss back with .item() every step")
    ap.add_argument("--graph", choices=["on", "off"], default="on",
                    help="replay the iteration's compute (sampler .. backward) as one CUDA graph (levels2fm_b200.graph.GraphedStep); "
                         "c4 then runs sphere_tracing in its synchronisation-free form (SDF.st_sync_free)")
    args = ap.parse_args()
    if args.workload in FORWARD_ONLY:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "forward-only workloads have no reference arm"}))
            return
        return run_forward_only(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from levels2fm_b200 import _C, ops, parallel, synthetic
    from levels2fm_b200.models.RadF import RadF
    from levels2fm_b200.models.Renderer import Renderer
    from levels2fm_b200.models.SDF import SDF

    _C.get()          # fail loudly if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    stdout_fd = None
    if world > 1:
        # stdout carries exactly ONE JSON line: NCCL prints its version banner on fd 1 from C, so fd 1 points at stderr until then
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device(dev))
    W = max(args.warmup, 3)

    wl = WORKLOADS[args.workload]
    strong = wl["scaling"] == "strong"
    if strong and wl["rays"] % world:
        raise SystemExit(f"{wl['rays']} rays do not split over {world} ranks")
    rays_rank = wl["rays"] // world if strong else wl["rays"]          # rays THIS rank renders per step
    rays_step = wl["rays"] if strong else wl["rays"] * world            # rays the whole job renders per step
    cams = wl["cams"]
    opt = workload_opt(args.workload, dev)
    half = float(opt.data.bound_max[0])
    torch.manual_seed(0)                                  # identical replicas on every rank
    sdf, rad, ren = SDF(opt).to(dev), RadF(opt).to(dev), Renderer(opt)
    synthetic.init_fields(sdf, rad, args.regime)
    if rad.dual_field and args.regime == "trained":
        with torch.no_grad():
            rad.embed_fn.embedder_obj.params.normal_(0.0, 0.05)
    params = list(sdf.parameters()) + list(rad.parameters())
    symmetric = args.symmetric == "on" or (args.symmetric == "auto" and world >= 8)
    bucket = parallel.GradBucket(params, symmetric=symmetric)
    H, Wd = opt.data.image_size
    if strong:          # every rank draws the iteration's full ray set with the SAME seed and keeps its contiguous slice
        c_all, r_all = synthetic.make_rays(cams, wl["rays"] // cams, half, H, Wd, seed=0)
        g_all = torch.rand(cams, wl["rays"] // cams, 3, generator=torch.Generator().manual_seed(1000))
        center_h, ray_h = parallel.shard_rays(c_all, r_all, rank, world)
        gt_h = g_all[:, rank * (g_all.shape[1] // world):(rank + 1) * (g_all.shape[1] // world)].contiguous()
    elif world > 1:
        # weak scaling, sharded the way SURVEY 8(e) shards an iteration: the job's batch is `world` camera sets (seed k = the batch a
        # single GPU would render as rank k) and every rank renders its contiguous 1/world slice of EVERY set -- same total work
        # (world x rays), but every rank sees the same mix of views, so the data-dependent sampler does not skew the ranks
        parts_c, parts_r, parts_g = [], [], []
        for k in range(world):
            c_k, r_k = synthetic.make_rays(cams, wl["rays"] // cams, half, H, Wd, seed=k)
            g_k = torch.rand(cams, wl["rays"] // cams, 3, generator=torch.Generator().manual_seed(1000 + k))
            cs, rs = parallel.shard_rays(c_k, r_k, rank, world)
            n = g_k.shape[1] // world
            parts_c.append(cs); parts_r.append(rs); parts_g.append(g_k[:, rank * n:(rank + 1) * n])
        center_h, ray_h, gt_h = torch.cat(parts_c, 0).contiguous(), torch.cat(parts_r, 0).contiguous(), torch.cat(parts_g, 0).contiguous()
    else:
        center_h, ray_h = synthetic.make_rays(cams, rays_rank // cams, half, H, Wd, seed=rank)
        gt_h = torch.rand(cams, rays_rank // cams, 3, generator=torch.Generator().manual_seed(1000 + rank))
    center_h, ray_h, gt_h = center_h.pin_memory(), ray_h.pin_memory(), gt_h.pin_memory()
    center, ray, gt = center_h.to(dev), ray_h.to(dev), gt_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # 256 MB > 126 MB L2
    # mean-normalised losses: with ray sharding every rank scales by the GLOBAL ray count so that the all-reduced SUM of the
    # per-rank gradients is the gradient of the global mean; rays_rank / rays_step = 1 / world
    loss_scale = rays_rank / rays_step if world > 1 else 1.0

    def compute(c, r, g):
        bucket.zero()
        out = ren.forward(opt, c, r, sdf, rad)
        loss = synthetic.render_loss_fused(out, g)
        if wl.get("trace"):      # depth-consistency term between the sphere-traced and the volume-rendered depth
            d_pred, _, _, _ = sdf.sphere_tracing(c, r, sdf)
            loss = loss + 1e-2 * (d_pred.view(out["depth_mlp"].shape[:2]) - out["depth_mlp"][..., 0].detach()).abs().mean()
        if loss_scale != 1.0:
            loss = loss * loss_scale
        loss.backward()
        return loss.detach(), out["sdfs_volume"].detach()

    use_graph = args.graph == "on"
    if wl.get("trace") and use_graph:
        sdf.st_sync_free = True      # sphere_tracing without its host read-back of the iteration count (same d_pred / finish_mask)
    graphed = None
    launches_per_step = None
    if use_graph:
        from levels2fm_b200.graph import GraphedStep
        try:
            ops.KLOG.reset()
            graphed = GraphedStep(compute, (center, ray, gt), warmup=W)
            launches_per_step = ops.KLOG.total() // (W + 1)          # W eager warm-up runs + the captured one
        except Exception as e:          # capture is an optimisation, never a requirement
            sys.stderr.write(f"CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly\n")
            graphed, use_graph = None, False

    def step(c, r, g):
        loss, sv = graphed(c, r, g) if graphed is not None else compute(c, r, g)
        bucket.allreduce()
        return loss, {"sdfs_volume": sv}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
        loss, out = step(center, ray, gt)
    n_samples = out["sdfs_volume"].shape[2]
    sync_all()

    # ---- device-resident timed region: K steps, per-step CUDA events, L2 flushed (untimed) between steps
    clocks = ClockSampler(local)
    clocks.start()
    ops.KLOG.reset()
    evs = []
    sync_all()
    for _ in range(args.steps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss, out = step(center, ray, gt)
        b.record()
        evs.append((a, b))
    sync_all()
    launches = ops.KLOG.total() if launches_per_step is None else launches_per_step * args.steps
    clk = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_local = sum(step_ms)
    t = torch.tensor([t_local], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = rays_step * args.steps / (total_ms / 1e3)

    # ---- end to end through the public API with HOST buffers: every step copies its inputs from pinned host memory (inside the
    #      timed region) and reads its loss back into a pinned slot; same L2 flush between steps and the same per-step CUDA events
    #      as the device-resident loop above, so `e2e` and `value` differ by exactly the copies.
    loss_host = torch.empty(args.steps, dtype=torch.float32).pin_memory()
    for k in range(3):          # untimed: first use of the pinned read-back slots and of the host-to-device staging blocks
        c = center_h.to(dev, non_blocking=True)
        loss, _ = step(c, ray_h.to(dev, non_blocking=True), gt_h.to(dev, non_blocking=True))
        loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
    sync_all()
    evs = []
    for k in range(args.steps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        c = center_h.to(dev, non_blocking=True)
        r = ray_h.to(dev, non_blocking=True)
        g = gt_h.to(dev, non_blocking=True)
        loss, _ = step(c, r, g)
        if args.e2e_sync_readback:
            loss_host[k] = loss.item()
        else:
            loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        b.record()
        evs.append((a, b))
    sync_all()
    assert bool(torch.isfinite(loss_host).all()), "e2e: non-finite loss read back"
    e2e_t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = rays_step * args.steps / (float(e2e_t.item()) / 1e3)
    h2d = center_h.numel() * 4 + ray_h.numel() * 4 + gt_h.numel() * 4

    # ---- per-kernel durations (separate pass, events around every launch) for the roofline of the dominant kernel
    n_prof = min(args.steps, 10)
    ops.KLOG.reset()
    ops.KLOG.timing = True
    coll_evs = []
    for _ in range(n_prof):
        flush.fill_(1.0)
        torch.cuda._sleep(int(8e6))      # ~4 ms of device-side spin: the host runs ahead, so every event pair brackets exactly its kernel
        compute(center, ray, gt)         # (eager: the events sit between the launches)
    torch.cuda.synchronize()
    ops.KLOG.timing = False
    launch_list = ops.KLOG.launches()
    per_name = {}
    for name, ms, units in launch_list:
        per_name.setdefault(name, []).append((ms, units))
    per_step_kernel_ms = {k: sum(m for m, _ in v) / n_prof for k, v in per_name.items()}
    dom = max(per_step_kernel_ms, key=per_step_kernel_ms.get)
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    G = GRID_BYTES_PER_EVAL
    # the LARGEST launch of the dominant kernel (a name may be launched several times per step with different sizes: the render's
    # field_backward next to the sphere-tracing track's): algorithmic bytes and duration of THAT launch only (DESIGN.md "Kernels"):
    #   field_backward: re-gather S*G + gradient scatter S*G + 52 B/sample of upstream grads and saved outputs
    #   field_forward : gather S*G + 28 B/sample of outputs
    big_units = max((u or 0) for _, u in per_name[dom])
    big_ms = statistics.mean(m for m, u in per_name[dom] if (u or 0) == big_units)
    S = big_units if big_units else rays_rank * n_samples
    alg = {"field_backward": S * (2 * G + 52), "field_forward": S * (G + 28)}.get(dom, S * G)
    achieved = alg / (big_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
    if os.path.exists(tr_path) and args.workload == "c2":
        traffic = json.load(open(tr_path)).get(dom)          # dram__bytes_read + dram__bytes_write per launch (ncu --set full)
    # what the load/store path allows for this kernel's scattered 8-byte accesses (measured with stand-alone gather / scatter kernels on
    # the same samples, profiles/l1tex_floor_r2.json): the ceiling the HBM-based `frac` can reach for this access pattern
    attainable = None
    fl_path = os.path.join(ROOT, "profiles", "l1tex_floor_r2.json")
    if os.path.exists(fl_path) and dom in ("field_backward", "field_forward"):
        fl = json.load(open(fl_path))
        floor_ms = S * (fl["gather_ns_per_sample"] + (fl["scatter_ns_per_sample"] if dom == "field_backward" else 0.0)) * 1e-6
        attainable = {"floor_ms": floor_ms, "frac_at_floor": alg / (floor_ms * 1e-3) / 1e9 / peak, "launch_ms_over_floor": big_ms / floor_ms,
                      "what": "L1TEX-bound gather" + (" + scatter" if dom == "field_backward" else "") + " of this launch's samples, nothing else (profiles/l1tex_floor_r2.json)"}
    # the sampler's up-sampling rounds that actually ran (rays still active), for the whole-step algorithmic bytes
    bpr = bytes_per_ray(wl, int(n_samples), n_trace=(3 * 10 if wl.get("trace") else 0))
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback", "attainable": attainable,
                "algorithmic_bytes_per_launch": alg, "launch_ms": big_ms, "launch_samples": S,
                "launches_of_this_kernel_per_step": len(per_name[dom]) // n_prof,
                "kernel_ms_per_step": per_step_kernel_ms,
                "whole_step": {"bytes_per_ray": bpr, "achieved": rays_rank * bpr / (ms_per_step * 1e-3) / 1e9,
                               "frac": rays_rank * bpr / (ms_per_step * 1e-3) / 1e9 / peak,
                               "note": "SURVEY 8(d) formula with k = 0 up-sampling rounds counted (a lower bound on the bytes)"}}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"],
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "name": args.workload, "rays_per_gpu": rays_rank, "rays_per_step": rays_step,
                           "samples_per_ray": int(n_samples), "regime": args.regime,
                           "l2": "flushed between timed steps (256 MB fill, untimed), in the device-resident loop AND in the e2e loop",
                           "cuda_graph": bool(use_graph),
                           "parallelism": f"ray-parallel dp{world}, one flat-bucket all-reduce per step" + (f" ({bucket.collective})" if world > 1 else "")
                                          + ("; every rank renders its 1/N slice of each of the N per-seed ray sets (SURVEY 8e sharding)" if world > 1 and not strong else "")},
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "roofline": roofline, "loss": float(loss_host[-1])}
    if world > 1:
        # where the multi-GPU step goes: the collective alone (events around bucket.allreduce on this rank, after a barrier so that
        # rank skew is excluded) and the skew itself (spread of the ranks' compute time)
        comp = torch.tensor([0.0], device=dev)
        coll_ms = []
        for _ in range(5):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            bucket.zero()
            out = ren.forward(opt, center, ray, sdf, rad)
            (synthetic.render_loss_fused(out, gt) * loss_scale).backward()
            b.record()
            torch.cuda.synchronize()
            comp += a.elapsed_time(b) / 5
            dist.barrier()
            torch.cuda.synchronize()
            a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a2.record()
            bucket.allreduce()
            b2.record()
            torch.cuda.synchronize()
            coll_ms.append(a2.elapsed_time(b2))
        cmax, cmin = comp.clone(), comp.clone()
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            nbytes = bucket.flat.numel() * 4
            line["multi_gpu"] = {"collective": bucket.collective, "bucket_bytes": nbytes, "collective_ms": statistics.median(coll_ms),
                                 "compute_ms_max_rank": float(cmax), "compute_ms_min_rank": float(cmin),
                                 "skew_ms": float(cmax - cmin),
                                 "nvlink_bytes_per_rank": int(2 * (world - 1) / world * nbytes)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        if world == 1 and not args.no_gpu_eager:
            try:
                del flush
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
                with torch.no_grad():
                    out = ren.forward(opt, center, ray, sdf, rad)
                    ours_loss = float(synthetic.render_loss(out, gt))
                    if wl.get("trace"):
                        d_pred, _, _, _ = sdf.sphere_tracing(center, ray, sdf)
                        ours_loss += float(1e-2 * (d_pred.view(out["depth_mlp"].shape[:2]) - out["depth_mlp"][..., 0]).abs().mean())
                del out
                eager = gpu_eager_baseline(args.workload, sdf, rad, center, ray, gt, ours_loss)
                eager["ours_over_eager"] = value / eager["value"]
                line["gpu_eager_baseline"] = eager
            except Exception as e:      # never lose the line over the baseline leg
                line["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
