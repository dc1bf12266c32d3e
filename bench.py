#!/usr/bin/env python
"""Benchmark of the Level-S2fM render hot path (BASELINE.json metric: rendered rays/sec, forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|uniform128] [--impl ours|reference]

One "step" = one optimisation iteration's render work on one batch of synthetic rays: depth sampling ->
fused field kernel (hash grid + SDF MLP + normals + radiance) -> compositing -> loss (10^3 L1 rgb + 10^2 eikonal)
-> backward (compositing backward + fused field backward); for N > 1 the single all-reduce of the flat gradient
bucket is inside the step.  Weak scaling: every rank renders its own 4096-ray batch.
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

RAYS_PER_GPU = 4096
WORKLOAD_TRACE = [False]      # set when the workload includes the per-iteration sphere_tracing call (c4)
GRID_BYTES_PER_EVAL = {16: 1024, 4: 256}     # L * 8 corners * F=2 * 4 B  


def workload_opt(name: str, device: str):
    from levels2fm_b200.config import default_opt
    over = {"SDF.arch.layers": [None, 64, 64, 64, 16], "RadF.arch.layers": [None, 64, 64, 3]}
    if name == "c2":
        over.update({"SDF.VolSDF.volsdf_sampling": True, "SDF.VolSDF.sample_intvs": 64, "SDF.VolSDF.final_sample_intvs": 64})
    elif name == "uniform128":
        over.update({"SDF.VolSDF.volsdf_sampling": False, "SDF.VolSDF.sample_intvs": 128})
    elif name == "c4":      # the reference's shipped DTU configuration + the sphere-tracing call CameraSet.render makes per iteration
        over = {"SDF.arch.layers": [None, 64, 16], "RadF.arch.layers": [None, 64, 64, 3],
                "SDF.VolSDF.volsdf_sampling": False, "SDF.VolSDF.sample_intvs": 128}
    else:
        raise ValueError(name)
    return default_opt("DTU", device=device, **over)


WORKLOAD_DESC = {
    "c2": "BASELINE configs[1]: 4096 rays, error-bounded sampler (64 coarse + 64 fine = 128 samples/ray), L=16 hash grid, "
          "SDF MLP 35-64-64-64-17, RadF 49-64-64-3, fused fwd+bwd, DTU bounds",
    "uniform128": "4096 rays, 128 uniform samples/ray, L=16 hash grid, SDF MLP 35-64-64-64-17, RadF 49-64-64-3, fused fwd+bwd, DTU bounds",
    "c4": "BASELINE configs[3] shape: 4096 rays, 128 uniform samples/ray, L=16 hash grid, shipped SDF MLP 35-64-17, RadF 49-64-64-3, "
          "one sphere_tracing call on all rays per iteration (pipelines/Camera.py:506), fused fwd+bwd, DTU bounds",
}


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


# ----------------------------------------------------------------------------- CPU reference arm
def oracle_setup(workload: str, n_rays: int, seed: int = 0):
    from levels2fm_b200 import synthetic
    from oracle import port
    cfg = port.SceneCfg(n_levels=16, sdf_layers=(None, 64, 16) if workload == "c4" else (None, 64, 64, 64, 16),
                        rad_layers=(None, 64, 64, 3), sample_intvs=64 if workload == "c2" else 128, final_sample_intvs=64,
                        volsdf_sampling=workload == "c2", iters_max_st=10)
    sdf_sd, rad_sd = port.random_state(cfg, seed=0, table_std=1e-4, generic_weights=False, sphere_bias=0.5)
    for sd in (sdf_sd, rad_sd):
        for k in sd:
            sd[k].requires_grad_(True)
    center, ray = synthetic.make_rays(1, n_rays, 1.0, 1200, 1600, seed=seed)
    gt = torch.rand(1, n_rays, 3, generator=torch.Generator().manual_seed(seed + 1))
    return cfg, sdf_sd, rad_sd, center, ray, gt


def oracle_step(cfg, sdf_sd, rad_sd, center, ray, gt):
    from levels2fm_b200 import synthetic
    from oracle import port
    for sd in (sdf_sd, rad_sd):
        for v in sd.values():
            v.grad = None
    out = port.render_forward(center, ray, sdf_sd, rad_sd, cfg)
    loss = synthetic.render_loss(out, gt)
    if len(cfg.sdf_layers) == 3 and not cfg.volsdf_sampling and cfg.sample_intvs == 128 and getattr(cfg, "_trace", True) and WORKLOAD_TRACE[0]:
        st = port.sphere_tracing(center, ray, sdf_sd, cfg)
        loss = loss + 1e-2 * (st["d_pred"] - out["depth_mlp"][..., 0].detach()).abs().mean()
    loss.backward()
    return float(loss.detach())


def cpu_baseline(workload: str, budget_s: float = 12.0, n_rays: int = 128):
    """The oracle port (reference algorithm restated in eager PyTorch, CPU) on a bounded sample of the workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    WORKLOAD_TRACE[0] = workload == "c4"
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_rays = 128
    WORKLOAD_TRACE[0] = args.workload == "c4"
    st = oracle_setup(args.workload, n_rays)
    for _ in range(max(args.warmup, 1)):
        oracle_step(*st)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle_step(*st)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = n_rays / (ms / 1e3)
    line = {"impl": "reference", "metric": "rendered rays/sec (fwd+bwd, 4096-ray batch)", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "rays_per_gpu": RAYS_PER_GPU, "regime": "init",
                       "sample": f"each step = {n_rays} rays of that workload (CPU, bounded)"},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} timed fwd+bwd steps of {n_rays} rays (oracle/port.py: the reference's algorithm "
                                       "restated in eager PyTorch on the host cores; tcnn/vren are not installable here)"},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regime", default="init", choices=["init", "trained"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-sync-readback", action="store_true", help="diagnostic: read the loss back with .item() every step")
    args = ap.parse_args()
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
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries exactly one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=torch.device(dev))
    W = max(args.warmup, 3)

    opt = workload_opt(args.workload, dev)
    torch.manual_seed(0)                                  # identical replicas on every rank
    sdf, rad, ren = SDF(opt).to(dev), RadF(opt).to(dev), Renderer(opt)
    synthetic.init_fields(sdf, rad, args.regime)
    params = list(sdf.parameters()) + list(rad.parameters())
    bucket = parallel.GradBucket(params)
    H, Wd = opt.data.image_size
    center_h, ray_h = synthetic.make_rays(1, RAYS_PER_GPU, 1.0, H, Wd, seed=rank)
    gt_h = torch.rand(1, RAYS_PER_GPU, 3, generator=torch.Generator().manual_seed(1000 + rank))
    center_h, ray_h, gt_h = center_h.pin_memory(), ray_h.pin_memory(), gt_h.pin_memory()
    center, ray, gt = center_h.to(dev), ray_h.to(dev), gt_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # 256 MB > 126 MB L2

    def step(c, r, g):
        bucket.zero()
        out = ren.forward(opt, c, r, sdf, rad)
        loss = synthetic.render_loss_fused(out, g)
        if args.workload == "c4":      # depth-consistency term between the sphere-traced and the volume-rendered depth
            d_pred, _, _, _ = sdf.sphere_tracing(c, r, sdf)
            loss = loss + 1e-2 * (d_pred - out["depth_mlp"][..., 0].detach()).abs().mean()
        loss.backward()
        bucket.allreduce()
        return loss, out

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
    launches = ops.KLOG.total()
    clk = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_local = sum(step_ms)
    t = torch.tensor([t_local], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * RAYS_PER_GPU * args.steps / (total_ms / 1e3)

    # ---- end to end through the public API with host buffers: every step copies its inputs from pinned host memory and reads its
    #      loss back to the host.  Both copies are asynchronous on the compute stream (the read-back lands in a pinned slot per step
    #      and is consumed after the loop), as a training loop that logs its loss would do it: no host stall inside the step.
    loss_host = torch.empty(args.steps, dtype=torch.float32).pin_memory()
    for k in range(3):          # untimed: first use of the pinned read-back slots and of the host-to-device staging blocks
        c = center_h.to(dev, non_blocking=True)
        loss, _ = step(c, ray_h.to(dev, non_blocking=True), gt_h.to(dev, non_blocking=True))
        loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        c = center_h.to(dev, non_blocking=True)
        r = ray_h.to(dev, non_blocking=True)
        g = gt_h.to(dev, non_blocking=True)
        loss, _ = step(c, r, g)
        if args.e2e_sync_readback:
            loss_host[k] = loss.item()
        else:
            loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
    sync_all()
    e2e_s = time.perf_counter() - t0
    assert bool(torch.isfinite(loss_host).all()), "e2e: non-finite loss read back"
    e2e_t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * RAYS_PER_GPU * args.steps / float(e2e_t.item())
    h2d = center_h.numel() * 4 + ray_h.numel() * 4 + gt_h.numel() * 4

    # ---- per-kernel durations (separate pass, events around every launch) for the roofline of the dominant kernel
    ops.KLOG.reset()
    ops.KLOG.timing = True
    for _ in range(min(args.steps, 10)):
        flush.fill_(1.0)
        torch.cuda._sleep(int(8e6))      # ~4 ms of device-side spin: the host runs ahead, so every event pair brackets exactly its kernel
        step(center, ray, gt)
    torch.cuda.synchronize()
    ops.KLOG.timing = False
    durs = {k: statistics.mean(v) for k, v in ops.KLOG.durations_ms().items()}
    counts = {k: len(v) // min(args.steps, 10) for k, v in ops.KLOG.durations_ms().items()}
    per_step_kernel_ms = {k: durs[k] * counts[k] for k in durs}
    dom = max(per_step_kernel_ms, key=per_step_kernel_ms.get)
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    G = GRID_BYTES_PER_EVAL[16]
    S = RAYS_PER_GPU * n_samples
    # algorithmic bytes of ONE launch of the dominant kernel (DESIGN.md "Kernels"):
    #   field_backward: re-gather S*G + gradient scatter S*G + 52 B/sample of upstream grads and saved outputs
    #   field_forward : gather S*G + 28 B/sample of outputs
    alg = {"field_backward": S * (2 * G + 52), "field_forward": S * (G + 28)}.get(dom, S * G)
    achieved = alg / (durs[dom] * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")
    if os.path.exists(tr_path) and args.workload == "c2":
        traffic = json.load(open(tr_path)).get(dom)          # dram__bytes_read + dram__bytes_write per launch (ncu --set full)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg, "launch_ms": durs[dom],
                "kernel_ms_per_step": per_step_kernel_ms}

    line = None
    if rank == 0:
        line = {"metric": "rendered rays/sec (fwd+bwd, 4096-ray batch)", "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_DESC[args.workload], "rays_per_gpu": RAYS_PER_GPU, "samples_per_ray": int(n_samples),
                           "regime": args.regime, "l2": "flushed between timed steps (256 MB fill, untimed)",
                           "parallelism": f"ray-parallel dp{world}, one flat-bucket all-reduce per step" + (f" ({bucket.collective})" if world > 1 else "")},
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "roofline": roofline, "loss": float(loss_host[-1])}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
